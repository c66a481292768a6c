# usage: bash scripts/profile_train.sh <tag>   -- ncu --set full (with source) of one epoch of train_epoch_kernel
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:train_epoch_kernel -s 1 -c 1 -f -o gpurun_out/prof_train_$TAG \
    python scripts/dev/dev_train_profile.py > gpurun_out/prof_train_$TAG.log 2>&1
ncu -i gpurun_out/prof_train_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_train_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_train_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_train_$TAG.csv 2>/dev/null
