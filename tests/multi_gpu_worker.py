"""Worker for tests/test_gpu_multi.py: launched by torch.distributed.run with one rank per GPU.
Checks the multi-GPU plumbing of the hot path on real devices (NCCL):
  1. a batch of chains sharded over ranks gives bit-identical end states to the same batch on one GPU;
  2. NestedSampler.run with chains sharded over the ranks: every rank ends with the identical evidence."""
import json
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    out_path = sys.argv[1]
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from nnest_b200 import NestedSampler, dist as nd
    from nnest_b200.engine import Engine
    from nnest_b200.likelihoods import Rosenbrock
    from helpers import load, state_dict_of

    # ---- 1. sharding invariance across real GPUs -------------------------------------------------------------
    g = load('mcmc_hard_rosen30.npz')
    d, n_per, steps = 30, 1024, 8
    n = n_per * world
    eng = Engine(local)
    eng.set_flow_from_state_dict(state_dict_of(g))
    eng.set_target(d, 0, [], t_scale=5.0, t_shift=0.0, prior_kind=1, prior_lo=-1.0, prior_hi=1.0)
    rng = np.random.RandomState(0)
    idx = rng.randint(0, g['active_u'].shape[0], size=n)
    u = np.ascontiguousarray(g['active_u'][idx].astype(np.float32).T)
    logl = g['active_logl'][idx]
    kw = dict(mode=0, loglstar=float(g['loglstar']), step_size=0.2, dynamic_step_size=False, seed=3)
    sl = slice(rank * n_per, (rank + 1) * n_per)
    st, _, _ = eng.mcmc_init(n_per, init_u=torch.from_numpy(np.ascontiguousarray(u[:, sl])).cuda(),
                             init_logl=torch.from_numpy(logl[sl]).cuda())
    eng.mcmc_run(st, steps, chain_offset=rank * n_per, **kw)
    gathered_x = nd.allgather_rows(st.x.t().contiguous())
    gathered_l = nd.allgather_rows(st.logl)
    ok_shard = True
    if rank == 0:
        full, _, _ = eng.mcmc_init(n, init_u=torch.from_numpy(u).cuda(), init_logl=torch.from_numpy(logl).cuda())
        eng.mcmc_run(full, steps, chain_offset=0, **kw)
        ok_shard = bool(torch.equal(gathered_x, full.x.t().contiguous()) and torch.equal(gathered_l, full.logl))

    # ---- 2. NestedSampler over the ranks ------------------------------------------------------------------------
    np.random.seed(11)           # same live points / chain starts on every rank (rank 0's are broadcast anyway)
    torch.manual_seed(11)
    s = NestedSampler(2, Rosenbrock(2), transform=lambda x: 5 * x, flow='nvp', num_live_points=400,
                      log_dir=os.path.join(os.path.dirname(out_path), 'logs'), log_level=logging.WARNING, seed=2)
    s.run(strategy=['mcmc'], mcmc_num_chains=256, mcmc_steps=10, train_iters=40)
    digest = torch.tensor([s.logz, s.h, float(s.niter), float(s.samples.sum())], dtype=torch.float64, device='cuda')
    parts = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(parts, digest)
    same = all(torch.equal(p, parts[0]) for p in parts)

    # ---- 3. the DEFAULT strategy (rejection_prior first), every rank with its OWN numpy stream ---------------------
    # (the rejection phase must gather the ranks' draws and agree on the switch to MCMC, as the reference
    # does at nested.py:295-298,366-378; otherwise the live sets diverge and the collectives dead-lock)
    np.random.seed(100 + rank)
    torch.manual_seed(100 + rank)
    s3 = NestedSampler(2, Rosenbrock(2), transform=lambda x: 5 * x, flow='nvp', num_live_points=300,
                       log_dir=os.path.join(os.path.dirname(out_path), 'logs3'), log_level=logging.WARNING, seed=4)
    s3.run(mcmc_num_chains=128, train_iters=30)
    digest3 = torch.tensor([s3.logz, s3.h, float(s3.niter), float(s3.samples.sum()), float(s3.total_calls >= 0)],
                           dtype=torch.float64, device='cuda')
    parts3 = [torch.empty_like(digest3) for _ in range(world)]
    dist.all_gather(parts3, digest3)
    same3 = all(torch.equal(p, parts3[0]) for p in parts3)
    if rank == 0:
        json.dump({'world': world, 'shard_ok': ok_shard, 'ranks_identical': bool(same), 'logz': float(s.logz),
                   'logzerr': float(s.logzerr), 'niter': int(s.niter), 'default_strategy_ranks_identical': bool(same3),
                   'default_strategy_logz': float(s3.logz), 'default_strategy_logzerr': float(s3.logzerr)},
                  open(out_path, 'w'))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
