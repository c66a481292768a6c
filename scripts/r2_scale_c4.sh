# usage: bash scripts/r2_scale_c4.sh N   -- c4 strong and weak scaling at N GPUs (the part of r2_scale.sh the README tables quote)
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sc in strong weak; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --scaling $sc > gpurun_out/r2_scale_n${N}_c4_$sc.json 2> gpurun_out/r2_scale_n${N}_c4_$sc.err || tail -c 800 gpurun_out/r2_scale_n${N}_c4_$sc.err
  python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/r2_scale_n${N}_c4_$sc.json') if l.startswith('{')][-1])
    print('N=$N $sc', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], 'launch_ms %.3f'%r['roofline']['launch_ms'], r['config']['kernel'], 'chains/gpu', r['config']['chains_per_gpu'], r['config']['shard_check']['hash'], 'e2e %.3e'%r['e2e']['value'], r.get('step_ms'))
except Exception as e: print('N=$N $sc ERR', e)
PY
done
