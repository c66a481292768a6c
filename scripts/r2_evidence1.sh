# round 2, GPU call 1: parity tests after the host-side changes, matched-settings evidence runs, C4 determinism
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_ev1_gpu.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_ev1_tests.log 2>&1
# matched small settings (the real reference runs the same on the CPU: scripts/ref_logz_seeds.py)
rm -f gpurun_out/r2_repo_logz.jsonl
(timeout 600 python scripts/repo_logz_seeds.py --seeds 0-9 --x_dim 10 --num_live_points 400 --mcmc_num_chains 400 --train_iters 50 --mcmc_steps 0 --tag S1 --out gpurun_out/r2_repo_logz.jsonl) > gpurun_out/r2_ev1_s1.log 2>&1
(timeout 600 python scripts/repo_logz_seeds.py --seeds 0-9 --x_dim 10 --num_live_points 400 --mcmc_num_chains 400 --train_iters 50 --mcmc_steps 10 --tag S3 --out gpurun_out/r2_repo_logz.jsonl) > gpurun_out/r2_ev1_s3.log 2>&1
# determinism: the same seed twice must give the same bits (4096 live d=30 variant, multi-CTA fit at batch 1024)
(timeout 600 python scripts/repo_logz_seeds.py --seeds 1-1 --x_dim 30 --num_live_points 4096 --mcmc_num_chains 4096 --train_iters 50 --batch_size 1024 --mcmc_steps 600 --strategy mcmc --tag C4s_rep_a --out gpurun_out/r2_repo_logz.jsonl) > gpurun_out/r2_ev1_rep_a.log 2>&1
(timeout 600 python scripts/repo_logz_seeds.py --seeds 1-6 --x_dim 30 --num_live_points 4096 --mcmc_num_chains 4096 --train_iters 50 --batch_size 1024 --mcmc_steps 600 --strategy mcmc --tag C4s --out gpurun_out/r2_repo_logz.jsonl) > gpurun_out/r2_ev1_c4s.log 2>&1
# under-fitted flow?  the reference's own training defaults (train_iters 500, batch 100)
(timeout 900 python scripts/repo_logz_seeds.py --seeds 1-3 --x_dim 30 --num_live_points 4096 --mcmc_num_chains 4096 --train_iters 500 --batch_size 100 --mcmc_steps 600 --strategy mcmc --tag C4s_fit500 --out gpurun_out/r2_repo_logz.jsonl) > gpurun_out/r2_ev1_c4s_fit.log 2>&1
tail -3 gpurun_out/r2_ev1_tests.log
wc -l gpurun_out/r2_repo_logz.jsonl
