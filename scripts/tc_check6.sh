cd $GRAFT_REPO_ROOT
NNB_TC_NPART=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
NNB_TC_NPART=1 NNB_NO_COOP=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
B="timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
for w in c4 c4 c5; do
  $B --workload $w 2>/dev/null | python profiles/benchline.py ${w}
done
NNB_TC_NPART=1 $B --workload c5 2>/dev/null | python profiles/benchline.py c5_npart1
