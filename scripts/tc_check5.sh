cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
NNB_TC_NPART=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
B="timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
for w in c4 c2 c3 c5; do
  $B --workload $w 2>/dev/null | python profiles/benchline.py ${w}_auto
done
NNB_TC_NPART=1 $B --workload c5 2>/dev/null | python profiles/benchline.py c5_npart1
NNB_TC_NPART=2 $B --workload c5 2>/dev/null | python profiles/benchline.py c5_npart2
NNB_TC_NPART=1 $B --chains 32768 2>/dev/null | python profiles/benchline.py c4_32k_npart1
NNB_TC_NPART=2 $B --chains 32768 2>/dev/null | python profiles/benchline.py c4_32k_npart2
