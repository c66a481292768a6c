cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/ev_tests.log 2>&1
(time timeout 300 python examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 8192 --mcmc_steps 1000 --log_dir /tmp/logs) > gpurun_out/ev_c5.log 2>&1
tail -3 gpurun_out/ev_c5.log
(time timeout 1000 python examples/nested/run.py --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --train_iters 50 --batch_size 8192 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 40000000 --log_dir /tmp/logs) > gpurun_out/ev_c4.log 2>&1
tail -5 gpurun_out/ev_c4.log
