cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
B="timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
for w in c4 c2 c3 c5; do
  $B --workload $w 2>/dev/null | python profiles/benchline.py $w
done
$B --chains 16384 2>/dev/null | python profiles/benchline.py c4_16k
NNB_TC_GENERIC=1 $B 2>/dev/null | python profiles/benchline.py c4_generic
./nnest_b200/lib/tc_latency | head -4
