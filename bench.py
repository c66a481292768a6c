#!/usr/bin/env python
"""bench.py -- proposals/sec of the latent-space MCMC hot path (flow inverse + log-Jacobian + prior +
likelihood + accept/reject), the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|c5] [--impl reference]

A "step" is one refill: `mcmc_steps` MCMC steps of every chain of the batch (one call of
Sampler._mcmc_sample, nnest/sampler.py:229) on synthetic inputs of the named shape.  Default workload
`c4` = BASELINE.json configs[3], the configuration the metric "at 1/2/4/8 B200" is quoted on
(Rosenbrock x_dim=30, 65536 chains, 150 steps, hard likelihood constraint, dynamic step size); chains
are sharded over ranks with a fixed number per GPU (weak scaling, no data-path collective).

Prints ONE JSON line (see the contract in the task statement).  `--impl reference` times the CPU
restatement of the reference's own path (oracle/, per-row Python loops like the reference) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (likelihood, d, chains per GPU, mcmc_steps, mode, transform scale, description)
    'c2': dict(like='himmelblau', d=2, chains=1024, mcmc_steps=10, mode='hard', ts=5.0, dynamic=True,
               desc='configs[1]: Himmelblau x_dim=2, 1024 chains x 10 steps, hard constraint'),
    'c3': dict(like='mixture', d=10, chains=2048, mcmc_steps=50, mode='hard', ts=10.0, dynamic=True,
               desc='configs[2]: GaussianMix x_dim=10, 2048 chains/GPU x 50 steps, hard constraint'),
    'c4': dict(like='rosenbrock', d=30, chains=65536, mcmc_steps=150, mode='hard', ts=5.0, dynamic=True,
               desc='configs[3]: Rosenbrock x_dim=30, 65536 chains/GPU x 150 steps, hard constraint, dynamic scale'),
    'c5': dict(like='gaussian', d=50, chains=32768, mcmc_steps=1000, mode='mh', ts=1.0, dynamic=False,
               desc='configs[4]: Gaussian(rho=0.99) x_dim=50, 32768 chains/GPU x 1000 steps, Metropolis-Hastings'),
}
HIDDEN, LAYERS, BLOCKS = 16, 1, 3      # reference defaults, nnest/sampler.py:37,42,43


def flow_flops_per_proposal(d):
    """Dense nn.Linear count of ONE flow inverse, 4*B*H*(2d + L*H)."""
    return 4 * BLOCKS * HIDDEN * (2 * d + LAYERS * HIDDEN)


def make_weights(d, seed=0):
    """Random-init SingleSpeedNVP weights (nn.Linear default U(-1/sqrt(fan_in), 1/sqrt(fan_in))), as the list
    of (weight (out,in), bias) in netG.state_dict() order: per block scale net then translate net."""
    rng = np.random.default_rng(seed)
    layers = []
    for _ in range(BLOCKS):
        for _net in ('scale', 'translate'):
            for o, i in [(HIDDEN, d)] + [(HIDDEN, HIDDEN)] * LAYERS + [(d, HIDDEN)]:
                b = 1.0 / np.sqrt(i)
                layers.append((rng.uniform(-b, b, size=(o, i)).astype(np.float32),
                               rng.uniform(-b, b, size=(o,)).astype(np.float32)))
    return layers


def flat_weights(layers):
    return np.concatenate([np.concatenate([w.ravel(), b.ravel()]) for w, b in layers])


def oracle_weights(layers, d):
    """(reference arm only) the same weights as an oracle.flow.NVPWeights"""
    from oracle import flow as oflow
    per_net = LAYERS + 2
    blocks = []
    for k in range(BLOCKS):
        base = k * 2 * per_net
        blocks.append({'scale': layers[base:base + per_net], 'translate': layers[base + per_net:base + 2 * per_net],
                       'const_scale': None})
    return oflow.NVPWeights(d, HIDDEN, LAYERS, BLOCKS, blocks)


def make_problem(wl, loglike_fn, seed=0):
    """Synthetic inputs shared by both arms: random-init flow of the named architecture, live points = best
    10% of uniform prior draws (Likelihood.uniform_sample, likelihoods.py:38-42), loglstar = min(active_logl),
    chain starts = active_u[randint] (nested.py:405-407).  loglike_fn(u (m,d) float64) -> (m,) float64."""
    d, n = wl['d'], wl['chains']
    rng = np.random.RandomState(seed)
    nlive = min(n, 16384)
    prob = dict(layers=make_weights(d, seed))
    if wl['mode'] == 'hard':
        u = rng.uniform(-1, 1, size=(nlive * 10, d))
        logl = np.asarray(loglike_fn(u), dtype=np.float64)
        order = np.argsort(-logl, kind='stable')[:nlive]
        active_u, active_logl = u[order], logl[order]
        idx = rng.randint(0, nlive, size=n)
        prob.update(init_u=active_u[idx], init_logl=active_logl[idx], loglstar=float(active_logl.min()))
    else:
        prob.update(init_z=(0.5 * rng.normal(size=(n, d))).astype(np.float32), loglstar=None)
    return prob


def oracle_like(wl):
    from oracle import likelihoods as olike
    d = wl['d']
    return {'rosenbrock': lambda: olike.Rosenbrock(d), 'himmelblau': lambda: olike.Himmelblau(2),
            'mixture': lambda: olike.GaussianMix(d), 'gaussian': lambda: olike.Gaussian(d, 0.99)}[wl['like']]()


def oracle_setup(wl, n_cpu):
    """(reference arm / cpu_baseline only) problem + target factory for the oracle port."""
    from oracle import likelihoods as olike, mcmc as omcmc
    d = wl['d']
    like = oracle_like(wl)
    prob = make_problem(dict(wl, chains=n_cpu), lambda u: like.batch(wl['ts'] * u))
    w = oracle_weights(prob['layers'], d)
    tr = (lambda x: wl['ts'] * x) if wl['mode'] == 'hard' else (lambda x: x * np.ones(d) + np.zeros(d))
    prior = olike.UniformPrior(d, -1, 1) if wl['mode'] == 'hard' else olike.UniformPrior(d, -5, 5)

    def run(steps_cpu):
        target = omcmc.Target(like, transform=tr, prior=prior, transform_prior=wl['mode'] != 'hard', rowwise=True)
        kw = dict(init_samples=prob['init_u'], init_loglikes=prob['init_logl'], loglstar=prob['loglstar'],
                  step_size=1 / d ** 0.5, dynamic_step_size=wl['dynamic']) if wl['mode'] == 'hard' else \
            dict(init_z=prob['init_z'], loglstar=None)
        omcmc.mcmc_sample(w, target, steps_cpu, omcmc.TorchNoise(), **kw)

    return run


LIKE_IDS = {'rosenbrock': (0, lambda d: []), 'himmelblau': (1, lambda d: []), 'gaussian': (2, lambda d: [0.99]),
            'mixture': (4, lambda d: [4.0, 1.0, 4.0, 0.4, 0.3, 0.2, 0.1])}


class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i] == 'Active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def run_reference(args, wl):
    """CPU arm: the oracle port of Sampler._mcmc_sample with the reference's per-row likelihood / prior loops
    and BLAS-threaded float32 matmuls, on a bounded sample (1024 chains) of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_cpu = min(1024, wl['chains'])
    steps_cpu = min(wl['mcmc_steps'], 20 if wl['d'] <= 30 else 5)
    run = oracle_setup(wl, n_cpu)

    def one_step():
        run(steps_cpu)

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    value = n_cpu * steps_cpu * args.steps / dt
    import torch
    cores = torch.get_num_threads()
    sample = '%d chains x %d mcmc steps per step (of %d x %d)' % (n_cpu, steps_cpu, wl['chains'], wl['mcmc_steps'])
    print(json.dumps({
        'impl': 'reference', 'metric': 'latent-space MCMC proposals/sec (flow+loglike)', 'value': value,
        'unit': 'proposals/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload + ': ' + wl['desc'], 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'proposals/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'proposals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def cpu_baseline_quick(wl, budget_s=15.0):
    """Oracle port (reference-style per-row loops) on rank 0 for ~budget_s seconds."""
    import torch
    n_cpu = min(1024, wl['chains'])
    run = oracle_setup(wl, n_cpu)
    steps_cpu = 5
    done, t0 = 0, time.perf_counter()
    while True:
        run(steps_cpu)
        done += n_cpu * steps_cpu
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {'value': done / dt, 'unit': 'proposals/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d chains, %d proposals in %.1f s (per-row prior/likelihood loops as in the reference)'
                      % (n_cpu, done, dt)}


def run_gpu(args, wl):
    import torch
    import torch.distributed as dist
    from nnest_b200 import build as nb_build
    from nnest_b200 import _lib as L

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    nb_build.build()
    from nnest_b200.engine import Engine

    d, n, S = wl['d'], wl['chains'], wl['mcmc_steps']
    mode = L.NNB_MODE_HARD if wl['mode'] == 'hard' else L.NNB_MODE_MH
    eng = Engine(local)
    like_id, like_params = LIKE_IDS[wl['like']][0], LIKE_IDS[wl['like']][1](d)
    if wl['mode'] == 'hard':
        eng.set_target(d, like_id, like_params, t_scale=wl['ts'], t_shift=0.0, prior_kind=L.NNB_PRIOR_BOX_U,
                       prior_lo=-1.0, prior_hi=1.0)
    else:
        eng.set_target(d, like_id, like_params, t_scale=1.0, t_shift=0.0, compute_f64=True,
                       prior_kind=L.NNB_PRIOR_BOX_V, prior_lo=-5.0, prior_hi=5.0)
    # identical on every rank; each rank runs its own global chain ids
    prob = make_problem(wl, lambda u: eng.loglike(torch.from_numpy(u).cuda()).cpu().numpy())
    eng.set_flow(flat_weights(prob['layers']), d, HIDDEN, LAYERS, BLOCKS, 0)
    chain_offset = rank * n
    step_size = 1 / d ** 0.5 if wl['mode'] == 'hard' else 0.0

    # ---- device-resident arm: inputs already in HBM ------------------------------------------------
    if wl['mode'] == 'hard':
        init_u_dev = torch.from_numpy(np.ascontiguousarray(prob['init_u'].astype(np.float32).T)).cuda()
        init_logl_dev = torch.from_numpy(prob['init_logl']).cuda()
        init_kw = dict(init_u=init_u_dev, init_logl=init_logl_dev)
    else:
        init_kw = dict(init_z=torch.from_numpy(np.ascontiguousarray(prob['init_z'].T)).cuda())
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')   # > 126 MB L2

    kernel_impl = {'auto': L.NNB_IMPL_AUTO, 'ffma': L.NNB_IMPL_FFMA, 'tcgen05': L.NNB_IMPL_TCGEN05}[args.kernel]
    run_ev = []

    def one_refill(it, timed=False):
        # device-resident arm: nothing is read back between refills, the calls only enqueue work (sync=False)
        st, _, _ = eng.mcmc_init(n, seed=args.seed, chain_offset=chain_offset, **init_kw)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        out = eng.mcmc_run(st, S, mode=mode, loglstar=prob['loglstar'], step_size=step_size,
                           dynamic_step_size=wl['dynamic'], seed=args.seed, chain_offset=chain_offset,
                           step_offset=it * S, impl=kernel_impl, sync=False)
        if timed:
            b.record()
            run_ev.append((a, b))
        return st, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for it in range(args.warmup):
        one_refill(it)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = eng.gpu_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for it in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not inside the event pair)
        ev[it][0].record()
        st, out = one_refill(args.warmup + it, timed=True)
        ev[it][1].record()
    barrier()
    last = eng.mcmc_result()                # counters of the last refill
    naccept, ncall = last['naccept'], last['ncall']
    launches = eng.gpu_launches - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    clk = clocks.stop() if rank == 0 else None
    proposals = n * S * args.steps * world
    value = proposals / (total_ms * 1e-3)

    # ---- end-to-end arm: host buffers in, host end states out, every step --------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    if wl['mode'] == 'hard':
        h_u, h_l = pin(prob['init_u'].astype(np.float32)), pin(prob['init_logl'])
        h2d = h_u.numel() * 4 + h_l.numel() * 8
    else:
        h_z = pin(prob['init_z'])
        h2d = h_z.numel() * 4
    h_first = torch.empty((d, n), dtype=torch.float32).pin_memory()
    h_last = torch.empty((d, n), dtype=torch.float32).pin_memory()
    h_logl = torch.empty((n,), dtype=torch.float64).pin_memory()
    d2h = (h_first.numel() + h_last.numel()) * 4 + h_logl.numel() * 8

    copy_stream = torch.cuda.Stream()

    def one_refill_e2e(it):
        if wl['mode'] == 'hard':
            kw = dict(init_u=h_u.cuda(non_blocking=True).t().contiguous(), init_logl=h_l.cuda(non_blocking=True))
        else:
            kw = dict(init_z=h_z.cuda(non_blocking=True).t().contiguous())
        st, _, _ = eng.mcmc_init(n, seed=args.seed, chain_offset=chain_offset, **kw)
        # the start points go back to the host on a side stream while the step kernel runs
        first = st.x.clone()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            h_first.copy_(first, non_blocking=True)
            first.record_stream(copy_stream)
        eng.mcmc_run(st, S, mode=mode, loglstar=prob['loglstar'], step_size=step_size,
                     dynamic_step_size=wl['dynamic'], seed=args.seed, chain_offset=chain_offset, step_offset=it * S,
                     impl=kernel_impl)
        h_last.copy_(st.x, non_blocking=True)
        h_logl.copy_(st.logl, non_blocking=True)
        torch.cuda.synchronize()

    one_refill_e2e(0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.steps):
        one_refill_e2e(it)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = proposals / (t.item() * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        # dominant kernel = the fused MCMC step kernel (mcmc_tc_kernel / mcmc_kernel): >99% of the step's device time
        # (profiles/).  Its launches are timed live with CUDA events around nnb_mcmc_run on the launching stream.
        run_ms = float(np.mean([a.elapsed_time(b) for a, b in run_ev]))
        kernel_launches = out['launches']
        flops = flow_flops_per_proposal(d)
        used_tc = out['impl'] == L.NNB_IMPL_TCGEN05
        achieved_tflops = n * S * flops / (run_ms * 1e-3) / 1e12           # algorithmic flops of one refill / its duration
        fp32_peak = 148 * 128 * 2 * ((clk or {}).get('sm_max_mhz') or 1965.0) * 1e6 / 1e12
        if used_tc and peaks.get('bf16_tflops'):
            peak, peak_src = peaks['bf16_tflops'] / 2.0, 'tf32 dense = 1/2 of the measured bf16 burst peak (MEASURED_PEAKS.json)'
        elif used_tc:
            peak, peak_src = 1590.0 / 2.0, 'tf32 dense = 1/2 of the fallback bf16 peak (B200_PROFILING.md)'
        else:
            peak, peak_src = fp32_peak, 'FP32 FMA pipe: 148 SMs x 128 lanes x 2 x max SM clock (not in MEASURED_PEAKS.json)'
        roofline = {
            'bound': 'tensor' if used_tc else 'fp32-fma', 'achieved': achieved_tflops, 'peak': peak, 'unit': 'TFLOP/s',
            'frac': achieved_tflops / peak, 'traffic': 9.38e6 * (n * S) / (65536.0 * 150) if used_tc else None,
            'kernel': 'mcmc_tc_kernel<MODE,NPART,D> (tcgen05 3xTF32)' if used_tc else 'mcmc_kernel<16,MODE> (FP32 FMA)',
            'launches_per_step': kernel_launches, 'launch_ms': run_ms / max(kernel_launches, 1),
            'algorithmic_flop_per_proposal': flops, 'peak_source': peak_src,
            'frac_of_fp32_fma_peak': achieved_tflops / fp32_peak,
            'note': 'algorithmic flops = dense nn.Linear count of one flow inverse per proposal, 4*B*H*(2d+L*H); the '
                    'tensor pipe executes them as 3xTF32 (x3) on mask-reduced operands (x0.6).  The kernel is bound by '
                    'the per-element work around the MMAs (Philox/Box-Muller, tanh/exp, hi/lo split, likelihood: FP32, ALU '
                    'and MUFU issue slots) and by the latency of nine dependent MMA round trips per step at 16 warps per '
                    'SM, not by the MMA rate: ncu (profiles/r1_final4_*) issue slots 48.5% busy, tensor pipe 9.9%, DRAM '
                    '0.04%; traffic = ncu dram bytes of one refill launch (c4) scaled to this size',
        }
        cpu = cpu_baseline_quick(wl) if world == 1 and not args.no_cpu_baseline else None
        print(json.dumps({
            'metric': 'latent-space MCMC proposals/sec (flow+loglike)', 'value': value, 'unit': 'proposals/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload + ': ' + wl['desc'], 'chains_per_gpu': n, 'x_dim': d,
                       'mcmc_steps': S, 'hidden_dim': HIDDEN, 'num_blocks': BLOCKS, 'num_layers': LAYERS,
                       'flow_weights': 'random init (nn.Linear default)', 'l2': 'flushed between timed steps',
                       'kernel': 'tcgen05' if out['impl'] == L.NNB_IMPL_TCGEN05 else 'ffma',
                       'accept_rate': naccept / float(n * S),
                       'loglike_calls_per_proposal': ncall / float(n * S)},
            'e2e': {'value': e2e_value, 'unit': 'proposals/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches,
            'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='c4', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--kernel', default='auto', choices=['auto', 'ffma', 'tcgen05'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--chains', type=int, default=0, help='development: override the chains per GPU of the workload')
    ap.add_argument('--mcmc-steps', type=int, default=0, help='development: override the MCMC steps per refill')
    ap.add_argument('--fixed-scale', action='store_true', help='development: dynamic_step_size=False')
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.chains or args.mcmc_steps or args.fixed_scale:
        wl.update(chains=args.chains or wl['chains'], mcmc_steps=args.mcmc_steps or wl['mcmc_steps'],
                  dynamic=wl['dynamic'] and not args.fixed_scale)
        wl['desc'] += ' [development override: %d chains x %d steps, dynamic=%s]' % (wl['chains'], wl['mcmc_steps'],
                                                                                      wl['dynamic'])
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == '__main__':
    main()
