# parity tests of the MCMC kernels + bench of the c4 workload for a few accept-phase noise splits
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q 2>&1 | tail -4
for jc in -1 2 4 5 6; do
  NNB_TC_JC=$jc timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python profiles/benchline.py jc$jc
done
