"""Command-line driver with the flags of the reference's examples/nested/run.py (:61-89), running on nnest_b200.

    python examples/nested/run.py --x_dim 2 --likelihood rosenbrock --flow nvp

Differences: --flow defaults to 'nvp' (the accelerated flow; the reference defaults to 'spline'), and
--batch_size / --seed are exposed.  Prints the evidence next to the analytic value where one is known."""
import argparse
import datetime
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')))

# uniform prior on the boxes of the reference's run.py:25-44; Rosenbrock by transfer-matrix quadrature
# (examples/nested/analytic_rosenbrock.py), the others by 2-D quadrature
ANALYTIC = {('rosenbrock', 2): -5.8041, ('rosenbrock', 3): -10.4770, ('rosenbrock', 10): -43.1084,
            ('rosenbrock', 30): -137.4875, ('himmelblau', 2): -5.5038, ('eggbox', 2): 235.895}


def main(args):
    from nnest_b200 import NestedSampler
    from nnest_b200.likelihoods import Himmelblau, Rosenbrock, Gaussian, Eggbox, GaussianShell, GaussianMix

    name = args.likelihood.lower()
    if name == 'himmelblau':
        like, transform = Himmelblau(args.x_dim), (lambda x: 5 * x)
    elif name == 'rosenbrock':
        like, transform = Rosenbrock(args.x_dim), (lambda x: 5 * x)
    elif name == 'gaussian':
        like, transform = Gaussian(args.x_dim, args.corr, lim=3), (lambda x: 3 * x)
    elif name == 'eggbox':
        like, transform = Eggbox(args.x_dim), (lambda x: x * 5 * np.pi)
    elif name == 'shell':
        like, transform = GaussianShell(args.x_dim), (lambda x: 5 * x)
    elif name == 'mixture':
        like, transform = GaussianMix(args.x_dim), (lambda x: 10 * x)
    else:
        raise ValueError('Likelihood not found')

    log_dir = os.path.join(args.log_dir, args.likelihood) + args.log_suffix
    if args.seed is not None:
        import torch
        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
    sampler = NestedSampler(like.x_dim, like, transform=transform, log_dir=log_dir,
                            num_live_points=args.num_live_points, hidden_dim=args.hidden_dim,
                            num_layers=args.num_layers, num_blocks=args.num_blocks, num_slow=args.num_slow,
                            use_gpu=True, scale=args.scale, flow=args.flow, batch_size=args.batch_size,
                            seed=args.seed or 0)
    start_time = time.time()
    sampler.run(train_iters=args.train_iters, mcmc_steps=args.mcmc_steps, volume_switch=args.switch,
                jitter=args.jitter, mcmc_num_chains=args.mcmc_num_chains,
                mcmc_dynamic_step_size=not args.mcmc_fixed_step_size, log_interval=args.log_interval,
                update_interval=args.update_interval, chain_stats=not args.no_chain_stats, max_iters=args.max_iters,
                strategy=args.strategy.split(',') if args.strategy else None)
    elapsed = time.time() - start_time
    if not sampler.single_or_primary_process:      # under torchrun: one rank reports
        return
    print('Run time %s' % datetime.timedelta(seconds=elapsed))
    ana = ANALYTIC.get((name, args.x_dim))
    if name == 'mixture':
        ana = -args.x_dim * np.log(20.0)
    out = dict(likelihood=name, x_dim=args.x_dim, gpus=sampler.mpi_size, num_live_points=args.num_live_points,
               mcmc_num_chains=args.mcmc_num_chains, logz=float(sampler.logz), logzerr=float(sampler.logzerr),
               h=float(sampler.h), niter=int(sampler.niter), ncall=int(sampler.total_calls), seconds=elapsed,
               analytic_logz=ana, sigma=None if ana is None else float((sampler.logz - ana) / sampler.logzerr))
    print(json.dumps(out))


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--x_dim', type=int, default=2, help="Dimensionality")
    parser.add_argument('--train_iters', type=int, default=2000, help="number of train iters")
    parser.add_argument('--mcmc_steps', type=int, default=0)
    parser.add_argument('--mcmc_num_chains', type=int, default=10)
    parser.add_argument('--num_live_points', type=int, default=1000)
    parser.add_argument('-mcmc_fixed_step_size', action='store_true')
    parser.add_argument('--switch', type=float, default=-1)
    parser.add_argument('--hidden_dim', type=int, default=16)
    parser.add_argument('--num_layers', type=int, default=1)
    parser.add_argument('-use_gpu', action='store_true')
    parser.add_argument('--flow', type=str, default='nvp')
    parser.add_argument('--num_blocks', type=int, default=3)
    parser.add_argument('--jitter', type=float, default=-1)
    parser.add_argument('--num_slow', type=int, default=0)
    parser.add_argument('--log_dir', type=str, default='logs')
    parser.add_argument('--likelihood', type=str, default='rosenbrock')
    parser.add_argument('--log_suffix', type=str, default='')
    parser.add_argument('--base_dist', type=str, default='')
    parser.add_argument('--scale', type=str, default='')
    parser.add_argument('--beta', type=float, default=8.0)
    parser.add_argument('--corr', type=float, default=0.99)
    parser.add_argument('--batch_size', type=int, default=100)
    parser.add_argument('--seed', type=int, default=None)
    parser.add_argument('--strategy', type=str, default='')
    parser.add_argument('--max_iters', type=int, default=1000000, help="NestedSampler.run(max_iters=...)")
    parser.add_argument('--log_interval', type=int, default=None, help="NestedSampler.run(log_interval=...)")
    parser.add_argument('--update_interval', type=int, default=None, help="NestedSampler.run(update_interval=...)")
    parser.add_argument('-no_chain_stats', action='store_true', help="skip the ESS/jump statistics at log lines")
    main(parser.parse_args())
