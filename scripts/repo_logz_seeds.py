#!/usr/bin/env python
"""Evidence of nnest_b200 on the B200 over several seeds at one setting (the counterpart of scripts/ref_logz_seeds.py,
which runs the REAL reference on the CPU at the same setting).  One JSON line per seed.

    python scripts/repo_logz_seeds.py --seeds 0-9 --x_dim 10 --num_live_points 400 --mcmc_num_chains 400 \
        --train_iters 50 --mcmc_steps 0 --out gpurun_out/repo_logz.jsonl
"""
import argparse
import json
import logging
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', default='0-9')
    ap.add_argument('--x_dim', type=int, default=10)
    ap.add_argument('--likelihood', default='rosenbrock')
    ap.add_argument('--num_live_points', type=int, default=400)
    ap.add_argument('--mcmc_num_chains', type=int, default=400)
    ap.add_argument('--mcmc_steps', type=int, default=0)
    ap.add_argument('--train_iters', type=int, default=50)
    ap.add_argument('--batch_size', type=int, default=100)
    ap.add_argument('--strategy', default='rejection_prior,mcmc')
    ap.add_argument('--flow', default='nvp')
    ap.add_argument('--max_iters', type=int, default=100000000)
    ap.add_argument('--out', default='')
    ap.add_argument('--tag', default='')
    args = ap.parse_args()

    import numpy as np
    import torch
    from nnest_b200 import NestedSampler
    from nnest_b200.likelihoods import Rosenbrock, Himmelblau, GaussianMix, Eggbox

    lo, _, hi = args.seeds.partition('-')
    seeds = range(int(lo), int(hi or lo) + 1)
    d = args.x_dim
    mk, ts = {'rosenbrock': (Rosenbrock, 5.0), 'himmelblau': (Himmelblau, 5.0), 'mixture': (GaussianMix, 10.0),
              'eggbox': (Eggbox, 5 * np.pi)}[args.likelihood]
    for seed in seeds:
        np.random.seed(seed)
        torch.manual_seed(seed)
        log_dir = tempfile.mkdtemp(prefix='repo_logz_')
        t0 = time.time()
        s = NestedSampler(d, mk(d), transform=lambda x: ts * x, log_dir=log_dir, num_live_points=args.num_live_points,
                          hidden_dim=16, num_layers=1, num_blocks=3, flow=args.flow, batch_size=args.batch_size,
                          log_level=logging.WARNING, seed=seed)
        s.run(strategy=args.strategy.split(','), train_iters=args.train_iters, mcmc_steps=args.mcmc_steps,
              mcmc_num_chains=args.mcmc_num_chains, max_iters=args.max_iters, log_interval=10 ** 9, chain_stats=False,
              diagnostics=True)
        rec = dict(impl='nnest_b200', seed=seed, x_dim=d, likelihood=args.likelihood, niter=int(s.niter),
                   ncall=int(s.total_calls), logz=float(s.logz), logzerr=float(s.logzerr), h=float(s.h),
                   logz_hex=float(s.logz).hex(), num_live_points=args.num_live_points,
                   mcmc_num_chains=args.mcmc_num_chains, mcmc_steps=args.mcmc_steps or 5 * d,
                   train_iters=args.train_iters, batch_size=args.batch_size, strategy=args.strategy, flow=args.flow,
                   wall_s=time.time() - t0, tag=args.tag)
        rl = np.array(getattr(s, 'refill_log', []) or np.zeros((0, 5)))
        fl = np.array(getattr(s.trainer, 'fit_log', []) or np.zeros((0, 5)))
        if len(rl):
            rec.update(n_refills=len(rl), acc_mean=float(rl[:, 2].mean()), usable_mean=float(rl[:, 3].mean()),
                       usable_min=float(rl[:, 3].min()), scale_last=float(rl[-1, 4]))
        if len(fl):
            rec.update(n_fits=len(fl), val_loss_last=float(fl[-1, 4]), best_epoch_mean=float(fl[:, 3].mean()),
                       fit_val_trace=[round(float(v), 4) for v in fl[:: max(1, len(fl) // 12), 4]])
        line = json.dumps(rec)
        print(line, flush=True)
        if args.out:
            with open(args.out, 'a') as f:
                f.write(line + '\n')


if __name__ == '__main__':
    main()
