#!/bin/bash
timeout 300 python scripts/dev/nn_time.py 2>&1 | grep "mean_nn"
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -3
bash scripts/r2_ns_full.sh 2>&1 | tail -2 | cut -c1-1300
