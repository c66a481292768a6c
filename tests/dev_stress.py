"""Development stress test: repeat identical launches and compare the results.
  * MCMC step kernel (deterministic: no floating-point atomics): every repetition must be BIT-IDENTICAL;
  * fused fitting kernel with many CTAs per mini-batch (float atomics: order varies): repetitions must agree to ~1e-6."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from nnest_b200 import _lib as L  # noqa: E402
from nnest_b200.engine import Engine  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
eng = Engine(0)

# ---- MCMC: c4 workload ---------------------------------------------------------------------------------------------
wl = bench.WORKLOADS['c4']
d, n, S = wl['d'], wl['chains'], wl['mcmc_steps']
eng.set_target(d, 0, [], t_scale=5.0, t_shift=0.0, prior_kind=L.NNB_PRIOR_BOX_U, prior_lo=-1.0, prior_hi=1.0)
prob = bench.make_problem(wl, lambda u: eng.loglike(torch.from_numpy(u).cuda()).cpu().numpy())
eng.set_flow(bench.flat_weights(prob['layers']), d, 16, 1, 3, 0)
u = torch.from_numpy(np.ascontiguousarray(prob['init_u'].astype(np.float32).T)).cuda()
logl = torch.from_numpy(prob['init_logl']).cuda()
ref = None
bad = 0
for r in range(reps):
    st, _, _ = eng.mcmc_init(n, init_u=u, init_logl=logl, seed=5)
    out = eng.mcmc_run(st, S, mode=L.NNB_MODE_HARD, loglstar=prob['loglstar'], step_size=1 / d ** 0.5,
                       dynamic_step_size=True, seed=5)
    cur = (st.x.clone(), st.z.clone(), st.logl.clone(), out['naccept'], out['ncall'], out['scale'])
    if ref is None:
        ref = cur
    else:
        same = torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and torch.equal(cur[2], ref[2]) \
            and cur[3:] == ref[3:]
        if not same:
            bad += 1
            print('MCMC repetition %d differs: naccept %d vs %d, scale %.9g vs %.9g, x mismatches %d' % (
                r, cur[3], ref[3], cur[5], ref[5], int((cur[0] != ref[0]).sum())))
print('mcmc: %d repetitions, %d differ' % (reps, bad))

# ---- fitting: 64 CTAs per mini-batch ------------------------------------------------------------------------------------
rng = np.random.RandomState(0)
x = torch.from_numpy(rng.uniform(-1, 1, size=(65536, d)).astype(np.float32)).cuda()
xv = x[:4096].contiguous()
from oracle import train as otrain  # noqa: E402  (development script)
P = otrain.net_floats(d, 16, 1) * 6
w0 = torch.from_numpy(bench.flat_weights(prob['layers'])).cuda()
assert w0.numel() == P
ref, worst = None, 0.0
for r in range(reps):
    w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    tl = []
    for ep in range(3):
        t_, vs, grid = eng.train_epoch((d, 16, 1, 3), w, m, v, ep * 8, x, xv, 8192, jitter=0.01, lr=1e-3,
                                       weight_decay=1e-6, seed=3, epoch=ep)
        tl.append(t_)
    cur = (w.clone(), m.clone(), tl)
    if ref is None:
        ref = cur
    else:
        dw = float((cur[0] - ref[0]).abs().max())
        worst = max(worst, dw)
        if dw > 1e-5 or abs(cur[2][-1] - ref[2][-1]) > 1e-4 * abs(ref[2][-1]):
            print('fit repetition %d deviates: max |dw| %.3e, loss %.8g vs %.8g' % (r, dw, cur[2][-1], ref[2][-1]))
print('fit: %d repetitions on %d CTAs, worst max |dw| %.3e' % (reps, grid, worst))
