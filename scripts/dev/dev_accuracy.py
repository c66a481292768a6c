"""Development aid (not a test): prints the error margins of the MCMC kernels against the reference goldens."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from helpers import load, state_dict_of, rel_err
from oracle import likelihoods as olike
from nnest_b200.engine import Engine
eng = Engine(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
HARD = {'mcmc_hard_rosen2.npz': (olike.Rosenbrock(2), 5.0), 'mcmc_hard_mix10.npz': (olike.GaussianMix(10), 10.0),
        'mcmc_hard_rosen30.npz': (olike.Rosenbrock(30), 5.0), 'mcmc_hard_himmel2_fixed.npz': (olike.Himmelblau(2), 5.0)}
for impl in (1, 2):
    for name, (like, ts) in HARD.items():
        g = load(name); d = int(g['d'])
        eng.set_flow_from_state_dict(state_dict_of(g))
        eng.set_target(d, like.like_id, like.params(), t_scale=ts, t_shift=0.0, prior_kind=1, prior_lo=-1.0, prior_hi=1.0)
        st, _, _ = eng.mcmc_init(int(g['chains']), init_u=dev(g['init_samples'].astype(np.float32).T), init_logl=dev(g['init_loglikes']))
        out = eng.mcmc_run(st, int(g['steps']), mode=0, loglstar=float(g['loglstar']), step_size=float(g['step_size']),
                           dynamic_step_size=bool(g['dynamic']), trace=True, replay=(dev(g['normals']), dev(g['uniforms'])), impl=impl)
        lat = out['trace_z'].permute(2, 0, 1).cpu().numpy(); smp = out['trace_x'].permute(2, 0, 1).cpu().numpy()
        same = np.all(np.any(lat[:, 1:] != lat[:, :-1], axis=2) == np.any(g['latent'][:, 1:] != g['latent'][:, :-1], axis=2), axis=1)
        print('impl %d %-28s flipped %d  rel_err z %.2e x %.2e' % (impl, name, (~same).sum(), rel_err(lat[same], g['latent'][same]), rel_err(smp[same], g['samples'][same])))
