# round 2: validate the restructured bench.py on one GPU (all workloads, reference arm, ns_c4)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python bench.py > gpurun_out/r2_b1_c4.json 2> gpurun_out/r2_b1_c4.err); tail -c 600 gpurun_out/r2_b1_c4.err
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_b1_ref.json 2> gpurun_out/r2_b1_ref.err)
for w in c2 c3 c5; do
  (timeout 600 python bench.py --workload $w --steps 10 --no-cpu-baseline > gpurun_out/r2_b1_$w.json 2> gpurun_out/r2_b1_$w.err); tail -c 400 gpurun_out/r2_b1_$w.err
done
(timeout 900 python bench.py --workload ns_c4 --steps 2 > gpurun_out/r2_b1_ns.json 2> gpurun_out/r2_b1_ns.err); tail -c 600 gpurun_out/r2_b1_ns.err
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_b1_tests.log 2>&1
tail -3 gpurun_out/r2_b1_tests.log
for f in gpurun_out/r2_b1_*.json; do echo $f; head -c 1500 $f; echo; done
