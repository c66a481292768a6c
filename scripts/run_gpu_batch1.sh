cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash scripts/profile_tc.sh r1b
(time timeout 200 python -m cProfile -s cumtime examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 8192 --mcmc_steps 1000 --log_dir /tmp/logs 2>&1 | grep -v "^\[" | head -60) > gpurun_out/c5_cprofile.log 2>&1
(time timeout 300 python examples/nested/run.py --x_dim 10 --likelihood rosenbrock --num_live_points 16384 --mcmc_num_chains 16384 --train_iters 50 --batch_size 2048 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 40000000 --log_dir /tmp/logs 2>&1 | tail -4) > gpurun_out/ev_r10.log 2>&1
(time timeout 600 python examples/nested/run.py --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --mcmc_steps 600 --train_iters 50 --batch_size 8192 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 40000000 --log_dir /tmp/logs 2>&1 | tail -4) > gpurun_out/ev_c4_600.log 2>&1
tail -3 gpurun_out/ev_r10.log gpurun_out/ev_c4_600.log
