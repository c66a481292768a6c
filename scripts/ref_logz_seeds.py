#!/usr/bin/env python
"""Evidence of the REAL reference (adammoss/nnest, CPU, flow='nvp') at a small matched setting, one seed per call.

TEST / EVIDENCE INFRASTRUCTURE (build container only: /root/reference does not exist on the GPU box).  The same
settings are run through nnest_b200 on the B200 by scripts/repo_logz_seeds.py; both distributions are committed in
profiles/r2_logz_matched.md.

    python scripts/ref_logz_seeds.py --seed 3 --x_dim 10 --num_live_points 400 --mcmc_num_chains 400 \
        --train_iters 50 --mcmc_steps 0 --out gpurun_out/ref_logz.jsonl
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
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--x_dim', type=int, default=10)
    ap.add_argument('--likelihood', default='rosenbrock')
    ap.add_argument('--num_live_points', type=int, default=400)
    ap.add_argument('--mcmc_num_chains', type=int, default=400)
    ap.add_argument('--mcmc_steps', type=int, default=0)
    ap.add_argument('--train_iters', type=int, default=50)
    ap.add_argument('--batch_size', type=int, default=100)
    ap.add_argument('--strategy', default='rejection_prior,mcmc')
    ap.add_argument('--threads', type=int, default=1)
    ap.add_argument('--out', default='')
    ap.add_argument('--tag', default='')
    args = ap.parse_args()

    import numpy as np
    import torch
    torch.set_num_threads(args.threads)
    from oracle.refload import load_reference
    nnest = load_reference()
    from nnest.likelihoods import Rosenbrock, Himmelblau, GaussianMix, Eggbox

    d = args.x_dim
    like, ts = {'rosenbrock': (Rosenbrock, 5.0), 'himmelblau': (Himmelblau, 5.0), 'mixture': (GaussianMix, 10.0),
                'eggbox': (Eggbox, 5 * np.pi)}[args.likelihood]
    like = like(d)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    log_dir = tempfile.mkdtemp(prefix='ref_logz_')
    t0 = time.time()
    s = nnest.NestedSampler(d, like, transform=lambda x: ts * x, log_dir=log_dir, num_live_points=args.num_live_points,
                            hidden_dim=16, num_layers=1, num_blocks=3, flow='nvp', use_gpu=False,
                            batch_size=args.batch_size, log_level=logging.WARNING)
    s.run(strategy=args.strategy.split(','), train_iters=args.train_iters, mcmc_steps=args.mcmc_steps,
          mcmc_num_chains=args.mcmc_num_chains)
    import csv
    with open(os.path.join(s.logs['results'], 'final.csv')) as f:
        rows = list(csv.reader(f))
    rec = dict(zip(rows[0], [float(v) for v in rows[1]]))
    rec.update(impl='reference', seed=args.seed, x_dim=d, likelihood=args.likelihood,
               num_live_points=args.num_live_points, mcmc_num_chains=args.mcmc_num_chains,
               mcmc_steps=args.mcmc_steps or 5 * d, train_iters=args.train_iters, batch_size=args.batch_size,
               strategy=args.strategy, wall_s=time.time() - t0, tag=args.tag)
    line = json.dumps(rec)
    print(line)
    if args.out:
        with open(args.out, 'a') as f:
            f.write(line + '\n')


if __name__ == '__main__':
    main()
