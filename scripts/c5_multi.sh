cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2
NG=${1:-8}
AVAIL=$(free -g | awk '/Mem:/ {print $7}')
CH=32768; if [ "$AVAIL" -lt $((45 * NG)) ]; then CH=8192; fi
echo "chains per GPU: $CH"
(time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains $CH --mcmc_steps 1000 --log_dir /tmp/logs 2>&1 | tail -3) > gpurun_out/ev_c5_${NG}gpu.log 2>&1
cat gpurun_out/ev_c5_${NG}gpu.log | grep -v "^$"
