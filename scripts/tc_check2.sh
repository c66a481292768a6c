cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q 2>&1 | tail -4
B="timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
for w in c4 c2 c3 c5; do
  $B --workload $w 2>/dev/null | python profiles/benchline.py fixed_$w
  NNB_TC_GENERIC=1 $B --workload $w 2>/dev/null | python profiles/benchline.py generic_$w
done
