#!/bin/bash
timeout 120 python scripts/dev/tc_timing.py 2>&1 | grep "kernel ms"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_api.py -x -q 2>&1 | tail -8
