cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "c4 2048 warp" "c4 8192 warp" "c4 8192 tcgen05" "c4 65536 tcgen05"; do
  set -- $cfg
  for fs in "" "--fixed-scale"; do
  (timeout 300 python bench.py --workload $1 --chains $2 --kernel $3 --steps 20 --no-cpu-baseline $fs > gpurun_out/r2_fx.json 2> gpurun_out/r2_fx.err) || tail -c 500 gpurun_out/r2_fx.err
  python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/r2_fx.json') if l.startswith('{')][-1])
    print('$cfg $fs', 'value %.3e'%r['value'], 'launch_ms %.3f'%r['roofline']['launch_ms'], r['config']['kernel'], 'acc %.3f'%r['config']['accept_rate'])
except Exception as e: print('$cfg ERR', e)
PY
  done
done
