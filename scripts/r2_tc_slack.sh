cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2_sl_tests.log 2>&1
tail -5 gpurun_out/r2_sl_tests.log
for cfg in "c4 65536 auto" "c4 32768 auto" "c4 16384 auto" "c4 8192 tcgen05" "c3 16384 auto" "c5 32768 auto"; do
  set -- $cfg
  (timeout 300 python bench.py --workload $1 --chains $2 --kernel $3 --steps 30 --no-cpu-baseline > gpurun_out/r2_sl_$1_$2_$3.json 2> gpurun_out/r2_sl_$1_$2_$3.err) || tail -c 500 gpurun_out/r2_sl_$1_$2_$3.err
  python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/r2_sl_$1_$2_$3.json') if l.startswith('{')][-1])
    print('$cfg', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], 'launch_ms %.3f'%r['roofline']['launch_ms'], r['config']['kernel'], 'acc %.3f'%r['config']['accept_rate'], 'e2e %.3e'%r['e2e']['value'], 'frac %.4f'%r['roofline']['frac'])
except Exception as e: print('$cfg ERR', e)
PY
done
