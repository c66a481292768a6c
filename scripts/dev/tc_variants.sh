#!/bin/bash
run() { timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$1', 'value %.4g' % j['value'], 'ms %.3f' % j['ms_per_step'], j.get('step_ms'))"; }
for i in 1 2 3 4 5 6; do run base; done
for i in 1 2 3; do NNB_BENCH_NOSLEEP=1 run nosleep; done
