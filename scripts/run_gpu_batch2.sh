cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_stats.py tests/test_gpu_api.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -5
(time timeout 300 python examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 8192 --mcmc_steps 1000 --log_dir /tmp/logs 2>&1 | tail -4) > gpurun_out/ev_c5_8k.log 2>&1
(time timeout 600 python examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 32768 --mcmc_steps 1000 --log_dir /tmp/logs 2>&1 | tail -4) > gpurun_out/ev_c5_32k.log 2>&1
tail -4 gpurun_out/ev_c5_8k.log gpurun_out/ev_c5_32k.log
for seed in 1 2 3; do
(timeout 300 python examples/nested/run.py --x_dim 30 --likelihood rosenbrock --num_live_points 4096 --mcmc_num_chains 4096 --mcmc_steps 600 --train_iters 50 --batch_size 1024 --seed $seed --strategy mcmc --log_interval 4000000 --max_iters 40000000 --log_dir /tmp/logs 2>&1 | tail -1) >> gpurun_out/ev_c4_4096.log 2>&1
done
cat gpurun_out/ev_c4_4096.log
