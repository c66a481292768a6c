# round 2: C5 end to end through MCMCSampler.run, full C4 run, C4 at longer chains / the reference's fit defaults, ns_c4
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 300 python examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 32768 --mcmc_steps 1000 --log_dir /tmp/logs) > gpurun_out/r2_ev2_c5_full.log 2>&1; tail -5 gpurun_out/r2_ev2_c5_full.log | cut -c 1-400
(time timeout 300 python examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 32768 --mcmc_steps 1000 --trace_thin 10 --log_dir /tmp/logs) > gpurun_out/r2_ev2_c5_thin.log 2>&1; tail -5 gpurun_out/r2_ev2_c5_thin.log | cut -c 1-400
(time timeout 600 python examples/nested/run.py --x_dim 30 --likelihood rosenbrock --num_live_points 65536 --mcmc_num_chains 65536 --train_iters 50 --batch_size 8192 --seed 1 --strategy mcmc --log_interval 4000000 --max_iters 40000000 --log_dir /tmp/logs) > gpurun_out/r2_ev2_c4.log 2>&1; tail -4 gpurun_out/r2_ev2_c4.log | cut -c 1-400
(timeout 900 python bench.py --workload ns_c4 --steps 2 > gpurun_out/r2_ev2_ns.json 2> gpurun_out/r2_ev2_ns.err); tail -1 gpurun_out/r2_ev2_ns.json | cut -c 1-900
(timeout 900 python scripts/repo_logz_seeds.py --seeds 1-3 --x_dim 30 --num_live_points 4096 --mcmc_num_chains 4096 --train_iters 500 --batch_size 100 --mcmc_steps 1500 --strategy mcmc --tag C4s_fit500_1500 --out gpurun_out/r2_repo_logz2.jsonl) > gpurun_out/r2_ev2_c4s.log 2>&1
grep logz gpurun_out/r2_repo_logz2.jsonl | cut -c 1-220
