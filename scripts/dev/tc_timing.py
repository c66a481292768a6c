"""Development probe: where does a step of mcmc_tc_kernel<0,1,30> spend its time?

Needs a library built with the stamps compiled in:
    NNB_EXTRA_NVCC_FLAGS=-DNNB_TC_TIMING NNB_LIB_DIR=lib_timing python -m nnest_b200.build
    NNB_LIB_DIR=lib_timing python scripts/dev/tc_timing.py          (on the GPU box)
Runs the c4 refill of bench.py (65 536 chains x 150 steps, fitted flow) and prints, per phase of a step, the median /
p10 / p90 over CTAs x tiles x steps of the clock64 differences between the stamps (cycles), and the spread of the step
start over the grid (globaltimer).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from nnest_b200 import _lib as L
    from nnest_b200.engine import Engine, flatten_state_dict
    name = 'c4'
    wl = bench.WORKLOADS[name]
    d, S, n = wl['d'], wl['mcmc_steps'], int(os.environ.get('NNB_N', wl['chains']))
    prob = bench.make_problem(name, wl, n, seed=0)
    flat, fd, fh, fl, fb, fflags = flatten_state_dict(prob['sd'], '')
    eng = Engine(0)
    like_id, like_params = bench.LIKE_IDS[wl['like']][0], bench.LIKE_IDS[wl['like']][1](d)
    eng.set_target(d, like_id, like_params, t_scale=wl['ts'], t_shift=0.0, prior_kind=L.NNB_PRIOR_BOX_U,
                   prior_lo=-1.0, prior_hi=1.0)
    eng.set_flow(flat, fd, fh, fl, fb, fflags)
    init_kw = dict(init_u=torch.from_numpy(np.ascontiguousarray(prob['init_u'][:n].astype(np.float32).T)).cuda(),
                   init_logl=torch.from_numpy(np.ascontiguousarray(prob['init_logl'][:n])).cuda())
    ms = []
    for it in range(6):
        st, _, _ = eng.mcmc_init(n, seed=0, **init_kw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.mcmc_run(st, S, mode=L.NNB_MODE_HARD, loglstar=prob['loglstar'], step_size=1 / d ** 0.5,
                     dynamic_step_size=True, seed=0, step_offset=it * S, impl=L.NNB_IMPL_TCGEN05, sync=False)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    print('%s JC=%s kernel ms:' % (os.environ.get('NNB_LIB_DIR', 'lib'), os.environ.get('NNB_TC_JC')), ['%.3f' % m for m in ms])
    lib = L.load()
    if not hasattr(lib, 'nnb_debug_tc_timing'):      # a library without the stamps: kernel time only
        return
    fn = lib.nnb_debug_tc_timing
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t]
    nct, ntile, nstep, nslot = 160, 4, 160, 10
    buf = np.zeros(nct * ntile * nstep * nslot, dtype=np.uint64)
    rc = fn(buf.ctypes.data, buf.size)
    assert rc == 0, rc
    t = buf.reshape(nct, ntile, nstep, nslot).astype(np.int64)
    grid = int(np.sum(t[:, 0, 0, 0] != 0))
    print('CTAs with stamps:', grid)
    t = t[:grid, :, :S]
    live = t[:, :, 0, 0] != 0                     # (cta, tile) pairs that exist
    T = t[live]                                   # (pairs, steps, slots)
    print('tiles:', T.shape[0])
    names = ['flow inverse (9 round trips)', 'accept + likelihood + state update', 'next noise (Philox)',
             'grid barrier wait (after noise)']
    cyc = lambda x: '%7.0f / %7.0f / %7.0f' % tuple(np.percentile(x, [10, 50, 90]))
    for k in range(4):
        dlt = T[:, 1:-1, k + 1] - T[:, 1:-1, k]
        print('%-40s p10/p50/p90 cycles: %s' % (names[k], cyc(dlt)))
    for name, a, b in [('  proposal z + scale * noise', 0, 6), ('  flow proper', 6, 1), ('  ratio test + likelihood', 1, 7),
                       ('  tile barrier (accept count)', 7, 8), ('  state update', 8, 2)]:
        print('%-40s p10/p50/p90 cycles: %s' % (name, cyc(T[:, 1:-1, b] - T[:, 1:-1, a])))
    step = T[:, 2:-1, 0] - T[:, 1:-2, 0]
    print('%-40s p10/p50/p90 cycles: %s' % ('whole step (start to start)', cyc(step)))
    gap = T[:, 2:-1, 0] - T[:, 1:-2, 4]
    print('%-40s p10/p50/p90 cycles: %s' % ('barrier release -> next step start', cyc(gap)))
    # who is last at the barrier?  spread of arrival (stamp 2, per CTA = max over its tiles) in globaltimer is not
    # available (clock64 is per SM); use globaltimer of the step start instead
    g = t[:, :, 1:S - 1, 5].astype(np.float64)
    g[~live] = np.nan
    spread = np.nanmax(g, axis=(0, 1)) - np.nanmin(g, axis=(0, 1))
    print('spread of step start over the grid (globaltimer ns): p10/p50/p90 %s' % cyc(spread))
    # per-tile-slot view: do the partial tiles (slot 3: 64 chains) finish the flow earlier?
    for j in range(ntile):
        m = t[:, j, 0, 0] != 0
        if m.any():
            dl = t[m, j, 1:S - 1, 1] - t[m, j, 1:S - 1, 0]
            ac = t[m, j, 1:S - 1, 2] - t[m, j, 1:S - 1, 1]
            print('tile slot %d: flow p50 %.0f, accept p50 %.0f' % (j, np.median(dl), np.median(ac)))
    # time from the LAST tile of the grid finishing its noise to the median release
    np.save(os.path.join(ROOT, 'gpurun_out', 'tc_timing_raw.npy'), t[:, :, :S].astype(np.int64))


if __name__ == '__main__':
    main()
