#!/usr/bin/env python
"""Data-parallel flow fitting (north_star: "allreduce of gradients for data-parallel flow training") measured against the
replicated fit the package uses (every rank fits the same live set; rank 0's weights are broadcast).

    torchrun --nproc-per-node N scripts/dp_train_measure.py [--n 65536 --d 30 --batch 8192 --epochs 20]

Both variants run the C4 retrain shape (65 536 live points, x_dim 30, 90 % training split) with the fused kernel:
  replicated : nnb_train_epoch, ONE launch per epoch (Adam inside the kernel, mini-batch over up to 64 CTAs)
  dp         : per mini-batch, every rank runs its 1/N share through the same kernel in gradient-only mode
               (nnb_train_args.grad_only), ncclAllReduce(SUM) of the flat gradient over NVLink, Adam on the flat
               parameter vector (torch.optim.Adam's update rule, fused with torch._foreach-free tensor ops)
Reports ms per epoch (CUDA events, max over ranks) and checks that both reach the same loss after the same steps."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def adam_step(w, g, m, v, step, lr, b1, b2, eps, wd):
    g = g + wd * w
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    w.addcdiv_(m, (v.sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=65536)
    ap.add_argument('--d', type=int, default=30)
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--epochs', type=int, default=20)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = dist.get_rank() if world > 1 else 0
    import bench
    from nnest_b200.engine import Engine
    eng = Engine(local)
    d, arch = args.d, (args.d, 16, 1, 3)
    rng = np.random.RandomState(0)
    x = torch.from_numpy((0.3 * rng.normal(size=(args.n, d))).astype(np.float32)).cuda()
    n_valid = args.n // 10
    x_valid, x_train = x[:n_valid].contiguous(), x[n_valid:].contiguous()
    n_train = x_train.shape[0]
    w0 = torch.from_numpy(bench.flat_weights(bench.make_weights(d, 0))).cuda()
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.999, 1e-8, 1e-6
    nsteps = (n_train + args.batch - 1) // args.batch
    res = {}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- replicated: one launch per epoch ---------------------------------------------------------------------------
    w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    perm = torch.arange(n_train, device='cuda')
    for ep in range(3):
        eng.train_epoch(arch, w, m, v, ep * nsteps, x_train, x_valid, args.batch, perm=perm, jitter=0.01, lr=lr,
                        weight_decay=wd, seed=1, epoch=ep)
    w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for ep in range(args.epochs):
        tl, vl, grid = eng.train_epoch(arch, w, m, v, ep * nsteps, x_train, x_valid, args.batch, perm=perm, jitter=0.01,
                                       lr=lr, weight_decay=wd, seed=1, epoch=ep)
    e1.record()
    sync()
    res['replicated_ms_per_epoch'] = e0.elapsed_time(e1) / args.epochs
    res['replicated_grid'] = grid
    res['replicated_val_nll'] = vl / n_valid

    # ---- data parallel: share of every mini-batch + all-reduce + Adam -----------------------------------------------------
    def dp_epoch(w, m, v, ep, timed_parts=None):
        for s in range(nsteps):
            lo, hi = s * args.batch, min(n_train, (s + 1) * args.batch)
            cnt = hi - lo
            a, b = lo + (cnt * rank) // world, lo + (cnt * (rank + 1)) // world
            g = torch.zeros_like(w)
            eng.train_epoch(arch, w, None, None, 0, x_train[a:b], None, b - a, jitter=0.01, seed=1 + 7919 * rank,
                            epoch=ep * nsteps + s, grad_out=g, grad_only=True, batch_total=cnt)
            if world > 1:
                dist.all_reduce(g)
            adam_step(w, g, m, v, ep * nsteps + s + 1, lr, b1, b2, eps, wd)

    w2, m2, v2 = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    for ep in range(2):
        dp_epoch(w2, m2, v2, ep)
    w2, m2, v2 = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    sync()
    e0.record()
    for ep in range(args.epochs):
        dp_epoch(w2, m2, v2, ep)
    e1.record()
    sync()
    res['dp_ms_per_epoch'] = e0.elapsed_time(e1) / args.epochs
    _, vl2, _ = eng.train_epoch(arch, w2, m2, v2, 0, None, x_valid, args.batch, do_train=False)
    res['dp_val_nll'] = vl2 / n_valid
    t = torch.tensor([res['replicated_ms_per_epoch'], res['dp_ms_per_epoch']], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res['replicated_ms_per_epoch'], res['dp_ms_per_epoch'] = t.tolist()
    res.update(world=world, n=args.n, d=d, batch=args.batch, steps_per_epoch=nsteps, epochs=args.epochs,
               dp_us_per_step=1e3 * res['dp_ms_per_epoch'] / nsteps,
               replicated_us_per_step=1e3 * res['replicated_ms_per_epoch'] / nsteps)
    if rank == 0:
        line = json.dumps(res)
        print(line)
        if args.out:
            with open(args.out, 'a') as f:
                f.write(line + '\n')
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
