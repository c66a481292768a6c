cd $GRAFT_REPO_ROOT
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 3 2>/dev/null | tail -1 | cut -c1-1500 > gpurun_out/bench_n$n.json
cat gpurun_out/bench_n$n.json | cut -c1-200
done
