cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_stress.py -x -q -m gpu 2>&1 | tail -3) > gpurun_out/r2_m2b_tests.log 2>&1; cat gpurun_out/r2_m2b_tests.log
rm -f gpurun_out/r2_dp_train.jsonl
timeout 300 python scripts/dp_train_measure.py --out gpurun_out/r2_dp_train.jsonl 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_train_measure.py --out gpurun_out/r2_dp_train.jsonl 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dp_train_measure.py --batch 32768 --out gpurun_out/r2_dp_train.jsonl 2>&1 | tail -1
timeout 300 python scripts/dp_train_measure.py --batch 32768 --out gpurun_out/r2_dp_train.jsonl 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2_m2b_bench_strong.json 2> gpurun_out/r2_m2b_bench_strong.err || tail -c 1500 gpurun_out/r2_m2b_bench_strong.err
python - <<PY
import json
r=json.loads([l for l in open('gpurun_out/r2_m2b_bench_strong.json') if l.startswith('{')][-1])
print('strong2', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], 'launch_ms %.3f'%r['roofline']['launch_ms'], 'e2e %.3e'%r['e2e']['value'])
PY
(timeout 600 python scripts/repo_logz_seeds.py --seeds 0-9 --x_dim 30 --num_live_points 400 --mcmc_num_chains 400 --train_iters 50 --mcmc_steps 0 --tag S4 --out gpurun_out/r2_repo_logz_s4.jsonl) > gpurun_out/r2_m2b_s4.log 2>&1
tail -3 gpurun_out/r2_m2b_s4.log | cut -c 1-200
