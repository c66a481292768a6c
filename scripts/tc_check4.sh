cd $GRAFT_REPO_ROOT
B="timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B 2>/dev/null | python profiles/benchline.py c4_npart2
NNB_TC_NPART=1 NNB_TC_NPART1_FIXED=1 $B 2>/dev/null | python profiles/benchline.py c4_npart1_fixed
NNB_TC_NPART=1 NNB_TC_NPART1_FIXED=1 $B --chains 16384 2>/dev/null | python profiles/benchline.py c4_16k_npart1_fixed
$B --workload c5 2>/dev/null | python profiles/benchline.py c5
$B --workload c5 2>/dev/null | python profiles/benchline.py c5_again
