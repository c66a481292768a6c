cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R="python examples/nested/run.py --log_dir /tmp/logs"
(time timeout 300 $R --x_dim 2 --likelihood rosenbrock --num_live_points 1000 --mcmc_num_chains 10 --train_iters 200 --seed 1 2>&1 | tail -1) > gpurun_out/ev2_c1.log 2>&1
(time timeout 300 $R --x_dim 2 --likelihood himmelblau --num_live_points 4096 --mcmc_num_chains 1024 --train_iters 200 --batch_size 512 --seed 1 2>&1 | tail -1) > gpurun_out/ev2_c2.log 2>&1
(time timeout 300 $R --x_dim 10 --likelihood mixture --num_live_points 16384 --mcmc_num_chains 16384 --train_iters 100 --batch_size 1024 --seed 1 --strategy mcmc 2>&1 | tail -1) > gpurun_out/ev2_c3.log 2>&1
(time timeout 300 $R --x_dim 2 --likelihood eggbox --num_live_points 4096 --mcmc_num_chains 1024 --train_iters 200 --batch_size 512 --seed 1 2>&1 | tail -1) > gpurun_out/ev2_egg.log 2>&1
cat gpurun_out/ev2_c1.log gpurun_out/ev2_c2.log gpurun_out/ev2_c3.log gpurun_out/ev2_egg.log | grep -v "^$\|user\|sys"
(timeout 300 python -m cProfile -s tottime examples/nested/run.py --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --train_iters 50 --batch_size 8192 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 1500000 --log_dir /tmp/logs 2>&1 | grep -v "^\[" | head -45) > gpurun_out/c4_cprofile.log 2>&1
(time timeout 900 $R --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --mcmc_steps 1500 --train_iters 50 --batch_size 8192 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 40000000 2>&1 | tail -1) > gpurun_out/ev_c4_1500.log 2>&1
cat gpurun_out/ev_c4_1500.log | grep -v "^$\|user\|sys"
