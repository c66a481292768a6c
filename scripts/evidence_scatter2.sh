cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R="python examples/nested/run.py --log_dir /tmp/logs"
for seed in 1 4; do
timeout 300 $R --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --train_iters 50 --batch_size 8192 --seed $seed --strategy mcmc --log_interval 1000000 --max_iters 40000000 > gpurun_out/c4_seed$seed.log 2>&1
tail -1 gpurun_out/c4_seed$seed.log
grep "Step \[" gpurun_out/c4_seed$seed.log | cut -c1-200
done
