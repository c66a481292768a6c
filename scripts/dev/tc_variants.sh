#!/bin/bash
# development A/B driver (edited per experiment): here the default bench line twice + the API tests
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('c4', 'value %.4g' % j['value'], 'ms %.3f' % j['ms_per_step'], 'e2e %.4g' % j['e2e']['value'], j['e2e']['ms_per_step_parts'], j.get('step_ms'))"; done
timeout 900 python -m pytest tests/test_gpu_api.py -x -q 2>&1 | tail -3
