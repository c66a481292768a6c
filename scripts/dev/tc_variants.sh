#!/bin/bash
for v in lib_big lib; do
for w in c4 c3; do NNB_LIB_DIR=$v timeout 300 python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$v $w', 'value %.4g' % j['value'], 'ms %.3f' % j['ms_per_step'], 'e2e %.4g' % j['e2e']['value'], 'kernel_ms', j['roofline'].get('launch_ms'))"; done
done
for w in c5 c2; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$w', 'value %.4g' % j['value'], 'ms %.3f' % j['ms_per_step'], 'e2e %.4g' % j['e2e']['value'], 'kernel_ms', j['roofline'].get('launch_ms'))"; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
