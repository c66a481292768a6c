cd $GRAFT_REPO_ROOT
B="timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B 2>/dev/null | python profiles/benchline.py base
$B --fixed-scale 2>/dev/null | python profiles/benchline.py fixed
$B --chains 32768 2>/dev/null | python profiles/benchline.py n32k
$B --chains 16384 2>/dev/null | python profiles/benchline.py n16k
$B --chains 131072 2>/dev/null | python profiles/benchline.py n131k
$B --chains 32768 --fixed-scale 2>/dev/null | python profiles/benchline.py n32k_fixed
NNB_TC_NPART=1 $B 2>/dev/null | python profiles/benchline.py npart1
$B --kernel ffma 2>/dev/null | python profiles/benchline.py ffma
