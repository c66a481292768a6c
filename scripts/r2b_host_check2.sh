# round-2 late check #2 (per-fit overheads of the flow fit, device-resident live set handed to the fit, consume v2):
# the API tests (every NestedSampler path), the trainer tests, then the full config-4 run with its wall-time split
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 170 python -m pytest tests/test_gpu_api.py tests/test_gpu_train.py -k "not mean_nn and not gradient and not two_epochs" -m gpu -x -q --durations=6 2>&1 | tail -25) | tee gpurun_out/r2c_tests.log
NNB_NS_ITERS=40000000 timeout 100 python bench.py --workload ns_c4 --steps 1 --warmup 0 > gpurun_out/r2c_ns_full.json 2> gpurun_out/r2c_ns_full.err; tail -c 400 gpurun_out/r2c_ns_full.err; tail -1 gpurun_out/r2c_ns_full.json | cut -c 1-1600
