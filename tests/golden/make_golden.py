"""Generate the golden fixtures in tests/golden/ from the REAL reference (adammoss/nnest).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box has no reference tree; tests there read the committed .npz files written here.

Fixtures
  flow_*.npz    netG.state_dict() + inputs + reference forward/inverse outputs
                (nnest/trainer.py:247-269 -> nnest/networks.py:24-42,289-309)
  like.npz      reference likelihood / prior values on fixed inputs, float32 and float64
                (nnest/likelihoods.py, nnest/priors.py, through the transforms of examples/nested/run.py:25-44)
  mcmc_*.npz    Sampler._mcmc_sample (nnest/sampler.py:229-463) replayed with recorded noise
  nested_*.npz  a recorded NestedSampler.run (nnest/nested.py:269-510): live set, every MCMC batch it
                consumed, and the final evidence / posterior arrays
"""
import logging
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle.refload import load_reference  # noqa: E402

nnest = load_reference()
from nnest.trainer import Trainer  # noqa: E402
from nnest import NestedSampler, MCMCSampler  # noqa: E402
from nnest import likelihoods as rl  # noqa: E402
from nnest.priors import UniformPrior  # noqa: E402

TMP = tempfile.mkdtemp(prefix='nnest_golden_')


def sd_arrays(netG):
    return {'sd/' + k: v.detach().cpu().numpy().copy() for k, v in netG.state_dict().items()}


def make_flow(tag, d, hidden=16, layers=1, blocks=3, scale='', n=64, perturb=0.0, seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    t = Trainer(d, hidden_dim=hidden, num_layers=layers, num_blocks=blocks, flow='nvp', scale=scale,
                log_dir=None, log_level=logging.WARNING)
    if perturb > 0:
        with torch.no_grad():
            for p in t.netG.parameters():
                p.add_(perturb * torch.randn_like(p))
    x = np.random.normal(size=(n, d)).astype(np.float32)
    z, ldz = t.forward(x, to_numpy=True)
    xr, ldx = t.inverse(z, to_numpy=True)
    zin = (1.5 * np.random.normal(size=(n, d))).astype(np.float32)
    xo, ldo = t.inverse(zin, to_numpy=True)
    out = dict(d=d, hidden=hidden, layers=layers, blocks=blocks, scale=scale,
               x=x, fwd_z=z, fwd_ld=ldz, inv_of_fwd_x=xr, inv_of_fwd_ld=ldx,
               zin=zin, inv_x=xo, inv_ld=ldo)
    out.update(sd_arrays(t.netG))
    np.savez_compressed(os.path.join(HERE, 'flow_%s.npz' % tag), **out)
    print('flow', tag, 'roundtrip err', np.abs(xr - x).max(), 'ld sum', np.abs(ldz + ldx).max())


def make_like():
    rng = np.random.RandomState(1)
    out = {}
    cases = {
        'rosenbrock2': (rl.Rosenbrock(2), 2, lambda x: 5 * x),
        'rosenbrock30': (rl.Rosenbrock(30), 30, lambda x: 5 * x),
        'himmelblau': (rl.Himmelblau(2), 2, lambda x: 5 * x),
        'gaussian10': (rl.Gaussian(10, 0.99, lim=3), 10, lambda x: 3 * x),
        'gaussian50': (rl.Gaussian(50, 0.99, lim=3), 50, lambda x: 3 * x),
        'eggbox': (rl.Eggbox(2), 2, lambda x: x * 5 * np.pi),
        'mixture10': (rl.GaussianMix(10), 10, lambda x: 10 * x),
        'mixture2': (rl.GaussianMix(2), 2, lambda x: 10 * x),
        'shell5': (rl.GaussianShell(5), 5, lambda x: 5 * x),
    }
    for name, (like, d, tr) in cases.items():
        u = rng.uniform(-1, 1, size=(48, d))
        if name.startswith('gaussian'):
            u = u * 0.2
        for dt in (np.float32, np.float64):
            ud = u.astype(dt)
            v = tr(ud)
            logl = like(v)
            key = '%s/%s' % (name, np.dtype(dt).name)
            out[key + '/u'] = ud
            out[key + '/v'] = np.asarray(v)
            out[key + '/logl'] = np.asarray(logl)
    pr = UniformPrior(4, -1, 1)
    xp = rng.uniform(-1.2, 1.2, size=(64, 4)).astype(np.float32)
    xp[0, 1] = 1.0
    xp[1, 2] = -1.0
    xp[2, 0] = np.float32(1.0000001)
    out['prior/x'] = xp
    out['prior/logp'] = np.array([pr(r) for r in xp], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'like.npz'), **out)
    print('like: %d arrays' % len(out))


class Replay(object):
    """Feeds recorded noise to the reference through torch.randn_like / torch.rand."""

    def __init__(self, normals, uniforms):
        self.normals, self.uniforms = list(normals), list(uniforms)
        self.i = self.j = 0

    def __enter__(self):
        self._rl, self._r = torch.randn_like, torch.rand

        def randn_like(t, *a, **k):
            v = torch.from_numpy(self.normals[self.i])
            self.i += 1
            assert v.shape == t.shape
            return v

        def rand(shape, *a, **k):
            v = torch.from_numpy(self.uniforms[self.j])
            self.j += 1
            assert tuple(v.shape) == tuple(shape)
            return v

        torch.randn_like, torch.rand = randn_like, rand
        return self

    def __exit__(self, *a):
        torch.randn_like, torch.rand = self._rl, self._r


def make_mcmc_hard(tag, like, d, tr, nlive, chains, steps, dynamic, train_iters=20, seed=0, hidden=16):
    torch.manual_seed(seed)
    np.random.seed(seed)
    s = NestedSampler(d, like, transform=tr, num_live_points=nlive, flow='nvp', hidden_dim=hidden,
                      log_dir=os.path.join(TMP, tag), log_level=logging.WARNING)
    prior = UniformPrior(d, -1, 1)
    u = prior.sample(nlive * 10)
    logl_all, _ = s.loglike(u)
    order = np.argsort(-logl_all)[:nlive]
    active_u, active_logl = u[order], logl_all[order]
    s.trainer.train(active_u, max_iters=train_iters, jitter=-1.0)
    loglstar = active_logl.min()
    idx = np.random.randint(0, nlive, size=chains)
    init_samples, init_loglikes = active_u[idx], active_logl[idx]
    normals = [np.random.normal(size=(chains, d)).astype(np.float32) for _ in range(steps)]
    uniforms = [np.random.uniform(size=(chains,)).astype(np.float32) for _ in range(steps)]
    step_size = 1 / d ** 0.5
    with Replay(normals, uniforms):
        samples, latent, derived, loglikes, scale, ncall = s._mcmc_sample(
            steps, init_samples=init_samples, init_loglikes=init_loglikes,
            init_derived=np.empty((chains, 0)), loglstar=loglstar, step_size=step_size,
            dynamic_step_size=dynamic, plot_trace=False)
    out = dict(d=d, nlive=nlive, chains=chains, steps=steps, dynamic=dynamic, loglstar=loglstar,
               step_size=step_size, init_samples=init_samples, init_loglikes=init_loglikes,
               normals=np.array(normals), uniforms=np.array(uniforms), samples=np.ascontiguousarray(samples),
               latent=np.ascontiguousarray(latent), loglikes=np.ascontiguousarray(loglikes),
               scale=float(scale), ncall=int(ncall), active_u=active_u, active_logl=active_logl)
    out.update(sd_arrays(s.trainer.netG))
    np.savez_compressed(os.path.join(HERE, 'mcmc_%s.npz' % tag), **out)
    acc = np.mean(np.any(samples[:, 1:] != samples[:, :-1], axis=2))
    print('mcmc', tag, 'scale', scale, 'ncall', ncall, 'move rate', acc)


def make_mcmc_mh(tag, d, corr, chains, steps, seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    like = rl.Gaussian(d, corr, lim=5)
    prior = UniformPrior(d, -5, 5)
    s = MCMCSampler(d, like, prior=prior, flow='nvp', log_dir=os.path.join(TMP, tag), log_level=logging.WARNING)
    cov = np.eye(d) + corr * (1 - np.eye(d))
    training = np.random.multivariate_normal(np.zeros(d), cov, size=400)
    mean, std = np.mean(training, axis=0), np.std(training, axis=0)
    s.transform = lambda x: x * std + mean                      # mcmc.py:107-111
    s.trainer.train((training - mean) / std, max_iters=20, jitter=0.01)
    z0 = (0.5 * np.random.normal(size=(chains, d))).astype(np.float32)
    s.trainer.get_prior_samples = lambda n, to_numpy=False: torch.from_numpy(z0)
    normals = [np.random.normal(size=(chains, d)).astype(np.float32) for _ in range(steps)]
    uniforms = [np.random.uniform(size=(chains,)).astype(np.float32) for _ in range(steps)]
    with Replay(normals, uniforms):
        samples, latent, derived, loglikes, scale, ncall = s._mcmc_sample(
            steps, num_chains=chains, loglstar=None, plot_trace=False)   # as mcmc.py:114
    out = dict(d=d, corr=corr, chains=chains, steps=steps, mean=mean, std=std, z0=z0, prior_min=-5.0, prior_max=5.0,
               normals=np.array(normals), uniforms=np.array(uniforms), samples=np.ascontiguousarray(samples),
               latent=np.ascontiguousarray(latent), loglikes=np.ascontiguousarray(loglikes),
               scale=float(scale), ncall=int(ncall))
    out.update(sd_arrays(s.trainer.netG))
    np.savez_compressed(os.path.join(HERE, 'mcmc_%s.npz' % tag), **out)
    acc = np.mean(np.any(samples[:, 1:] != samples[:, :-1], axis=2))
    print('mcmc', tag, 'scale', scale, 'ncall', ncall, 'move rate', acc)


def make_nested(tag, like, d, tr, nlive, chains, seed=0):
    """Record a full reference run (strategy mcmc only, fixed step) and every batch it consumed."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    s = NestedSampler(d, like, transform=tr, num_live_points=nlive, flow='nvp',
                      log_dir=os.path.join(TMP, tag), log_level=logging.WARNING)
    rec = {'batches_first': [], 'batches_last': [], 'batches_logl': [], 'loglstar': [], 'idx': []}
    orig = s._mcmc_sample
    orig_prior_sample = s.sample_prior
    orig_randint = np.random.randint

    def sample_prior(n):
        u = orig_prior_sample(n)
        rec['active_u0'] = u.copy()
        return u

    s.sample_prior = sample_prior

    def rec_mcmc(*a, **k):
        out = orig(*a, **dict(k, plot_trace=False))
        samples, latent, derived, loglikes, scale, nc = out
        rec['batches_first'].append(samples[:, 0, :].copy())
        rec['batches_last'].append(samples[:, -1, :].copy())
        rec['batches_logl'].append(loglikes[:, -1].copy())
        rec['loglstar'].append(k['loglstar'])
        rec['init_samples_last'] = k['init_samples']
        return out

    s._mcmc_sample = rec_mcmc

    def randint(low=0, high=None, size=None):
        r = orig_randint(low=low, high=high, size=size)
        if size == chains:
            rec['idx'].append(r.copy())
        return r

    np.random.randint = randint
    s.trainer.train = lambda *a, **k: None         # untrained flow: bookkeeping is what is recorded
    try:
        s.run(strategy=['mcmc'], mcmc_num_chains=chains, mcmc_dynamic_step_size=False, mcmc_steps=5)
    finally:
        np.random.randint = orig_randint
    active_u0 = rec['active_u0']
    logl0, _ = s.loglike(active_u0)
    out = dict(d=d, nlive=nlive, chains=chains, active_u0=active_u0, active_logl0=logl0,
               batches_first=np.array(rec['batches_first']), batches_last=np.array(rec['batches_last']),
               batches_logl=np.array(rec['batches_logl']), loglstar=np.array(rec['loglstar']),
               idx=np.array(rec['idx']),
               logz=s.logz, samples=s.samples, weights=s.weights, loglikes=s.loglikes)
    import csv
    with open(os.path.join(s.logs['results'], 'final.csv')) as f:
        rows = list(csv.reader(f))
    out['final_header'] = np.array(rows[0])
    out['final_row'] = np.array([float(v) for v in rows[1]])
    with open(os.path.join(s.logs['chains'], 'chain.txt')) as f:
        out['chain_txt_head'] = np.array(f.read().splitlines()[:5])
    np.savez_compressed(os.path.join(HERE, 'nested_%s.npz' % tag), **out)
    print('nested', tag, 'logz', s.logz, 'niter', rows[1][0], 'batches', len(rec['batches_logl']))


if __name__ == '__main__':
    try:
        for d in (2, 3, 4, 5):                                    # dims of tests/test_flows.py:56-72
            make_flow('d%d' % d, d, seed=d)
        make_flow('d10_p', 10, perturb=0.3, seed=10)
        make_flow('d30_p', 30, perturb=0.2, seed=30)
        make_flow('d50_p', 50, perturb=0.15, seed=50)
        make_flow('d7_h32_l2_b5', 7, hidden=32, layers=2, blocks=5, perturb=0.2, seed=7)
        make_flow('d6_translate', 6, scale='translate', perturb=0.3, seed=6)
        make_flow('d6_constant', 6, scale='constant', perturb=0.3, seed=16)
        make_like()
        make_mcmc_hard('hard_rosen2', rl.Rosenbrock(2), 2, lambda x: 5 * x, 200, 32, 12, dynamic=True)
        make_mcmc_hard('hard_himmel2_fixed', rl.Himmelblau(2), 2, lambda x: 5 * x, 200, 32, 10, dynamic=False, seed=1)
        make_mcmc_hard('hard_mix10', rl.GaussianMix(10), 10, lambda x: 10 * x, 300, 32, 20, dynamic=True, seed=2)
        make_mcmc_hard('hard_rosen30', rl.Rosenbrock(30), 30, lambda x: 5 * x, 300, 16, 20, dynamic=True, seed=3)
        make_mcmc_hard('hard_eggbox2', rl.Eggbox(2), 2, lambda x: x * 5 * np.pi, 200, 32, 10, dynamic=True, seed=4)
        make_mcmc_mh('mh_gauss8', 8, 0.9, 24, 15, seed=5)
        make_mcmc_mh('mh_gauss50', 50, 0.99, 8, 8, seed=6)
        make_nested('rosen2', rl.Rosenbrock(2), 2, lambda x: 5 * x, 60, 8, seed=7)
    finally:
        shutil.rmtree(TMP, ignore_errors=True)
