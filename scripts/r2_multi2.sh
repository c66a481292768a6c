# 2 GPUs: NCCL test of the sharded hot path + the strong / weak scaled bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_m2_gpus.txt
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s -m gpu 2>&1 | grep -E "MULTI-GPU RESULT|passed|failed|Error" | tail -5) > gpurun_out/r2_m2_test.log 2>&1
cat gpurun_out/r2_m2_test.log
for sc in strong weak; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --scaling $sc > gpurun_out/r2_m2_bench_$sc.json 2> gpurun_out/r2_m2_bench_$sc.err || tail -c 1500 gpurun_out/r2_m2_bench_$sc.err
python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/r2_m2_bench_$sc.json') if l.startswith('{')][-1])
    print('$sc', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], 'launch_ms %.3f'%r['roofline']['launch_ms'], r['config']['kernel'], r['config']['chains_per_gpu'], r['config']['shard_check']['hash'], r['config']['collectives_per_step'], 'e2e %.3e'%r['e2e']['value'], r['e2e']['ms_per_step_parts'], (r.get('ns_loop') or {}).get('ms_per_refill_parts'))
except Exception as e: print('$sc ERR', e)
PY
done
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2_m2_bench_n1.json 2>/dev/null
python - <<PY
import json
r=json.loads([l for l in open('gpurun_out/r2_m2_bench_n1.json') if l.startswith('{')][-1])
print('n1', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], r['config']['shard_check']['hash'], 'e2e %.3e'%r['e2e']['value'])
PY
# the reference's CLI under torchrun (every rank builds the sampler; rank 0 alone owns the run directory)
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 examples/nested/run.py --x_dim 10 --likelihood mixture --num_live_points 16384 --mcmc_num_chains 4096 --train_iters 100 --batch_size 1024 --seed 1 --strategy mcmc --log_dir /tmp/logs2) > gpurun_out/r2_m2_c3_run.log 2>&1; tail -6 gpurun_out/r2_m2_c3_run.log | cut -c 1-400; ls /tmp/logs2 | head
