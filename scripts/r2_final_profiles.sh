# round 2: the profiles kept under profiles/ (one GPU): launch list of the bench command, ncu --set full of the step kernel
# (tcgen05 on c4, 16-lane on the N = 8 shard of c4) and of the fitting kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_r2_final.log 2>&1
bash scripts/profile_kernel.sh r2_final_tc mcmc_tc_kernel
bash scripts/profile_kernel.sh r2_final_warp mcmc_warp_kernel --chains 4096
ncu --set full --clock-control none --import-source on -k regex:train_epoch_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_final_train \
    python scripts/dev/dev_train_profile.py > gpurun_out/prof_r2_final_train.log 2>&1
ncu -i gpurun_out/prof_r2_final_train.ncu-rep --page raw --csv > gpurun_out/raw_r2_final_train.csv 2>/dev/null
tail -3 gpurun_out/prof_r2_final_train.log | cut -c 1-300
python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -c 300 gpurun_out/r2_final_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_final_bench_ref.json
