#!/bin/bash
# development A/B driver (edited per experiment; run on the GPU box from the repo root).  Last use: start-delay patterns of the
# tensor-core step kernel (NNB_TC_DELAYS) and the nearest-neighbour kernel's query tiles / splits (NNB_NN_QT, NNB_NN_SPLITS).
for dl in "0,700,0,700" "0,0,0,0" "0,600,0,600"; do
  NNB_TC_DELAYS=$dl timeout 120 python scripts/dev/tc_timing.py 2>&1 | grep "kernel ms" | sed "s/^/delays $dl: /"
done
for qt in 1 2 4; do NNB_NN_ONE=1 NNB_NN_QT=$qt timeout 100 python scripts/dev/nn_time.py 2>&1 | grep "mean_nn"; done
