#!/bin/bash
for sl in 0 1; do
  NNB_TC_SLACK=$sl timeout 120 python scripts/dev/tc_timing.py 2>&1 | grep "kernel ms" | sed "s/^/slack $sl: /"
done
NNB_TC_SLACK=1 NNB_TC_DELAYS=0,0,0,0 timeout 120 python scripts/dev/tc_timing.py 2>&1 | grep "kernel ms" | sed "s/^/slack 1 nodelay: /"
NNB_TC_SLACK=1 NNB_TC_DELAYS=0,400,0,400 timeout 120 python scripts/dev/tc_timing.py 2>&1 | grep "kernel ms" | sed "s/^/slack 1 d400: /"
NNB_LIB_DIR=lib_timing timeout 120 python scripts/dev/tc_timing.py 2>&1 | grep -v Warn
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q 2>&1 | tail -3
