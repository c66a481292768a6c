# the full config-4 nested-sampling run through bench.py --workload ns_c4 (all iterations), with the wall-time split
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NNB_NS_ITERS=40000000 timeout 900 python bench.py --workload ns_c4 --steps 1 --warmup 0 > gpurun_out/r2_ns_full.json 2> gpurun_out/r2_ns_full.err; tail -c 300 gpurun_out/r2_ns_full.err; tail -1 gpurun_out/r2_ns_full.json | cut -c 1-1500
