# round-2 late check of the host-side changes (queued epochs, chain writer, bookkeeping): the GPU tests that exercise them,
# then the full config-4 nested-sampling run with its wall-time split
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_train.py tests/test_gpu_stress.py "tests/test_gpu_api.py::test_nested_run_bookkeeping_bit_exact_and_layout" "tests/test_gpu_api.py::test_run_diagnostics_are_recorded" "tests/test_gpu_api.py::test_trainer_injection_b1" "tests/test_gpu_api.py::test_training_reduces_loss_and_updates_device_weights" -m gpu -x -q 2>&1 | tail -15) | tee gpurun_out/r2b_tests.log
NNB_NS_ITERS=40000000 timeout 150 python bench.py --workload ns_c4 --steps 1 --warmup 0 > gpurun_out/r2b_ns_full.json 2> gpurun_out/r2b_ns_full.err; tail -c 400 gpurun_out/r2b_ns_full.err; tail -1 gpurun_out/r2b_ns_full.json | cut -c 1-1600
