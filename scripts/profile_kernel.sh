# usage (on the GPU box, from the repo root): bash scripts/profile_kernel.sh <tag> <kernel regex> <bench args...>
# one ncu --set full capture (with source) of one kernel of a bench.py command
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/prof_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_$TAG.csv 2>/dev/null
tail -2 gpurun_out/prof_$TAG.log | head -c 600
