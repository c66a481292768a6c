# usage (on the GPU box, from the repo root): bash scripts/profile_tc.sh <tag>
# one ncu --set full capture (with source) of the fused MCMC step kernel on the c4 workload + the launch list
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mcmc_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_$TAG.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
