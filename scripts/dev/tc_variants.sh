#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in c4 c3 c5; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$w', 'value %.4g' % j['value'], 'ms %.3f' % j['ms_per_step'], 'e2e %.4g' % j['e2e']['value'], 'kernel_ms', j['roofline'].get('launch_ms'), 'frac', j['roofline']['frac'])"; done
