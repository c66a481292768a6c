# usage: bash scripts/r2_scale.sh N     (on a box with N GPUs): the bench at N ranks, strong and weak scaling, + named configs
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {  # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 "$@" > gpurun_out/r2_scale_n${N}_$tag.json 2> gpurun_out/r2_scale_n${N}_$tag.err || tail -c 800 gpurun_out/r2_scale_n${N}_$tag.err
  python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/r2_scale_n${N}_$tag.json') if l.startswith('{')][-1])
    print('N=$N $tag', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], 'launch_ms %.3f'%r['roofline']['launch_ms'], r['config']['kernel'], 'chains/gpu', r['config']['chains_per_gpu'], r['config']['shard_check']['hash'], 'e2e %.3e'%r['e2e']['value'], (r.get('ns_loop') or {}).get('ms_per_refill_parts'))
except Exception as e: print('N=$N $tag ERR', e)
PY
}
run c4_strong --scaling strong
run c4_weak --scaling weak
run c3_strong --workload c3 --scaling strong
run c5_strong --workload c5 --scaling strong --steps 5
if [ "$N" = "8" ]; then
  (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 32768 --mcmc_steps 1000 --trace_thin 10 --log_dir /tmp/logs) > gpurun_out/r2_scale_n8_c5_run.log 2>&1; tail -5 gpurun_out/r2_scale_n8_c5_run.log | cut -c 1-500
  (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 examples/nested/run.py --x_dim 10 --likelihood mixture --num_live_points 16384 --mcmc_num_chains 2048 --train_iters 100 --batch_size 1024 --seed 1 --strategy mcmc --log_dir /tmp/logs) > gpurun_out/r2_scale_n8_c3_run.log 2>&1; tail -4 gpurun_out/r2_scale_n8_c3_run.log | cut -c 1-500
fi
