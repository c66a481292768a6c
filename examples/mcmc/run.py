"""Latent-space MCMC driver (BASELINE.json configs[4]): MCMCSampler.run (nnest/mcmc.py:79-126) on the correlated
Gaussian of examples/nested/run.py:34-36, many concurrent chains, Metropolis-Hastings inside the fused CUDA kernel.

    python examples/mcmc/run.py --x_dim 50 --corr 0.99 --mcmc_num_chains 32768 --mcmc_steps 1000

The reference ships only a stale notebook for this workload (examples/mcmc/example.ipynb); this script drives the
same API.  It prints the posterior moments of the chains next to the analytic ones (mean 0, covariance
(1-rho) I + rho 11^T truncated by the +-lim box, negligible at lim = 5 sigma).  With --thin k only every k-th state
of the trace is used for the moments (the full trace stays on the device: 404 B per proposal at d = 50)."""
import argparse
import json
import logging
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')))


def main(args):
    import torch
    import torch.distributed as dist
    world, rank = 1, 0
    if 'LOCAL_RANK' in os.environ:        # one process per GPU under torchrun: every rank runs its own shard of chains
        local = int(os.environ['LOCAL_RANK'])
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        rank, world = dist.get_rank(), dist.get_world_size()
    from nnest_b200 import MCMCSampler
    from nnest_b200.likelihoods import Gaussian
    from nnest_b200.priors import UniformPrior

    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    d, rho = args.x_dim, args.corr
    cov = (1 - rho) * np.eye(d) + rho * np.ones((d, d))
    training = np.random.multivariate_normal(np.zeros(d), cov, size=args.training_samples)
    sampler = MCMCSampler(d, Gaussian(d, rho, lim=args.lim), prior=UniformPrior(d, -args.lim, args.lim), flow='nvp',
                          hidden_dim=args.hidden_dim, num_blocks=args.num_blocks, num_layers=args.num_layers,
                          batch_size=args.batch_size, log_dir=os.path.join(args.log_dir, 'gaussian'),
                          log_level=logging.INFO if rank == 0 else logging.WARNING, seed=args.seed)
    t0 = time.time()
    sampler.run(args.mcmc_steps, args.mcmc_num_chains, training, stats_interval=None, train_iters=args.train_iters,
                thin=args.trace_thin)
    elapsed = time.time() - t0
    burn = (args.mcmc_steps // 2) // args.trace_thin
    tail = sampler.samples[:, burn::max(1, args.thin // args.trace_thin), :d]
    flat = torch.from_numpy(np.ascontiguousarray(tail, dtype=np.float64).reshape(-1, d)).cuda()
    # moments over the chains of ALL ranks: sums are all-reduced (no sample ever leaves its GPU's host)
    stats = torch.cat([flat.sum(0), (flat.T @ flat).reshape(-1),
                       torch.tensor([flat.shape[0], float(sampler.total_accepted), float(sampler.total_rejected),
                                     float(sampler.total_calls)], dtype=torch.float64, device='cuda')])
    t_max = torch.tensor([elapsed], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(stats)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    stats = stats.cpu().numpy()
    cnt, acc, rej, ncall = stats[d + d * d:]
    mean = stats[:d] / cnt
    c = stats[d:d + d * d].reshape(d, d) / cnt - np.outer(mean, mean)
    n_chains = args.mcmc_num_chains * world
    out = dict(x_dim=d, corr=rho, gpus=world, chains=n_chains, chains_per_gpu=args.mcmc_num_chains,
               steps=args.mcmc_steps, seconds=float(t_max.item()), proposals=n_chains * args.mcmc_steps,
               ncall=int(ncall), acceptance=acc / max(1.0, acc + rej),
               max_abs_mean=float(np.abs(mean).max()), mean_sigma_over_sqrt_chains=float(1.0 / np.sqrt(n_chains)),
               var_mean=float(np.diag(c).mean()), var_expected=1.0,
               offdiag_mean=float((c.sum() - np.trace(c)) / (d * (d - 1))), offdiag_expected=rho,
               max_abs_cov_err=float(np.abs(c - cov).max()))
    if rank != 0:
        dist.destroy_process_group()
        return
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--x_dim', type=int, default=50)
    ap.add_argument('--corr', type=float, default=0.99)
    ap.add_argument('--lim', type=float, default=5.0)
    ap.add_argument('--mcmc_steps', type=int, default=1000)
    ap.add_argument('--mcmc_num_chains', type=int, default=32768)
    ap.add_argument('--training_samples', type=int, default=20000)
    ap.add_argument('--train_iters', type=int, default=200)
    ap.add_argument('--batch_size', type=int, default=1000)
    ap.add_argument('--hidden_dim', type=int, default=16)
    ap.add_argument('--num_blocks', type=int, default=3)
    ap.add_argument('--num_layers', type=int, default=1)
    ap.add_argument('--thin', type=int, default=10, help='stride of the states used for the printed moments')
    ap.add_argument('--trace_thin', type=int, default=1,
                    help='MCMCSampler.run(thin=...): rows of the trace brought to the host (1 = every state, as the reference)')
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--log_dir', type=str, default='logs')
    main(ap.parse_args())
