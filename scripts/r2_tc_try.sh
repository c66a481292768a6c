cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu 2>&1 | tail -4) > gpurun_out/r2_try_tests.log 2>&1
tail -3 gpurun_out/r2_try_tests.log
for fl in 0 1; do
for cfg in "c4 65536 auto" "c4 32768 auto" "c3 16384 auto"; do
  set -- $cfg
  (NNB_TC_FLAGS=$fl timeout 300 python bench.py --workload $1 --chains $2 --kernel $3 --steps 30 --no-cpu-baseline > gpurun_out/r2_try.json 2> gpurun_out/r2_try.err) || tail -c 500 gpurun_out/r2_try.err
  python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/r2_try.json') if l.startswith('{')][-1])
    print('flags=$fl $cfg', 'value %.3e'%r['value'], 'ms %.3f'%r['ms_per_step'], 'launch_ms %.3f'%r['roofline']['launch_ms'], r['config']['kernel'], 'acc %.3f'%r['config']['accept_rate'])
except Exception as e: print('$cfg ERR', e)
PY
done
done
(NNB_TC_FLAGS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2)
