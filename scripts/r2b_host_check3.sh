# round-2 late check #3: the spline-flow tests (autograd fit path of the Trainer after the fit-loop changes) and the
# default bench line (its ns_loop block runs the new live-point replacement)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 70 python -m pytest tests/test_gpu_spline.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -8) | tee gpurun_out/r2d_tests.log
timeout 100 python bench.py > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; tail -c 300 gpurun_out/r2d_bench_default.err; tail -1 gpurun_out/r2d_bench_default.json | cut -c 1-2500
