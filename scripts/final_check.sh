cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/final_tests.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) | tee gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c 1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -1 gpurun_out/bench_reference.json | cut -c 1-300
