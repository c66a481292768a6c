cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
B="timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline"
for w in c4 c4 c5 c3 c2; do
  $B --workload $w 2>/dev/null | python profiles/benchline.py ${w}
done
