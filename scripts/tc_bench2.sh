cd $GRAFT_REPO_ROOT
B="timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
for w in c2 c3; do
  $B --workload $w 2>/dev/null | python profiles/benchline.py ${w}_tc
  $B --workload $w --kernel ffma 2>/dev/null | python profiles/benchline.py ${w}_ffma
done
$B --chains 16384 2>/dev/null | python profiles/benchline.py c4_16k_auto
NNB_TC_NPART=1 $B --chains 16384 2>/dev/null | python profiles/benchline.py c4_16k_npart1
$B --chains 16384 --kernel ffma 2>/dev/null | python profiles/benchline.py c4_16k_ffma
$B --chains 4096 2>/dev/null | python profiles/benchline.py c4_4k_tc
$B --chains 4096 --kernel ffma 2>/dev/null | python profiles/benchline.py c4_4k_ffma
