cd $GRAFT_REPO_ROOT
R="python examples/nested/run.py --log_dir /tmp/logs"
timeout 200 $R --x_dim 10 --likelihood rosenbrock --num_live_points 16384 --mcmc_num_chains 16384 --train_iters 50 --batch_size 2048 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 40000000 2>&1 | tail -1
timeout 200 $R --x_dim 2 --likelihood himmelblau --num_live_points 4096 --mcmc_num_chains 1024 --train_iters 200 --batch_size 512 --seed 2 2>&1 | tail -1
for seed in 2 3; do
timeout 300 $R --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --train_iters 50 --batch_size 8192 --seed $seed --strategy mcmc --log_interval 4000000 --max_iters 40000000 2>&1 | tail -1
done
