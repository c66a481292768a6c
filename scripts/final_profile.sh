cd $GRAFT_REPO_ROOT
B="timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
for w in c2 c3; do
  NNB_TC_NPART=1 $B --workload $w 2>/dev/null | python profiles/benchline.py ${w}_npart1
  NNB_TC_NPART=2 $B --workload $w 2>/dev/null | python profiles/benchline.py ${w}_npart2
done
bash scripts/profile_tc.sh r1e
