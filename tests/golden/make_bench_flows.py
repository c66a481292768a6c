"""Fitted flows + synthetic live points for bench.py, recorded from the REAL reference (adammoss/nnest).

Run in the build container only (needs /root/reference):   python tests/golden/make_bench_flows.py

SURVEY.md section 8(d): "live points u ~ U[-1,1]^d drawn by UniformPrior.sample, keep the best nlive of nlive/fraction
by likelihood (Likelihood.uniform_sample, fraction 0.1) so the constraint is non-trivial; loglstar = min(active_logl);
flow = reference Trainer(flow='nvp') fitted for a fixed number of epochs on those points, its state_dict loaded into
both paths; chains start at active_u[randint]".  bench.py (both arms) reads the committed bench_<workload>.npz.

  bench_c2.npz  Himmelblau d=2      bench_c3.npz  GaussianMix d=10      bench_c4.npz  Rosenbrock d=30   (hard constraint)
  bench_c5.npz  Gaussian(rho=0.99) d=50, flow fitted to standardised posterior draws (MCMCSampler.run, mcmc.py:107-112)
"""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle.refload import load_reference  # noqa: E402

nnest = load_reference()
from nnest.trainer import Trainer  # noqa: E402
from nnest import likelihoods as rl  # noqa: E402
from nnest.priors import UniformPrior  # noqa: E402
from oracle import likelihoods as olike  # noqa: E402  (vectorised evaluation of 10 x nlive prior draws)

NLIVE, FRACTION, EPOCHS = 1024, 0.1, 30


def sd_arrays(netG):
    return {'sd/' + k: v.detach().cpu().numpy().copy() for k, v in netG.state_dict().items()}


def hard(tag, like, olk, d, ts):
    np.random.seed(0)
    torch.manual_seed(0)
    prior = UniformPrior(d, -1, 1)
    u = prior.sample(int(NLIVE / FRACTION))                       # likelihoods.py:38-42 on the unit box
    logl = np.asarray(olk.batch(ts * u), dtype=np.float64)
    order = np.argsort(-logl, kind='stable')[:NLIVE]
    active_u, active_logl = u[order], logl[order]
    chk = np.array([like(ts * r) for r in active_u[:8]])          # the reference's own per-row evaluation
    assert np.allclose(chk, active_logl[:8], rtol=1e-12, atol=1e-12), (chk, active_logl[:8])
    t = Trainer(d, hidden_dim=16, num_layers=1, num_blocks=3, flow='nvp', learning_rate=0.001, log_dir=None,
                log_level=logging.WARNING)
    t.train(active_u, max_iters=EPOCHS, jitter=-1.0)
    out = dict(d=d, ts=ts, nlive=NLIVE, epochs=EPOCHS, active_u=active_u.astype(np.float32),
               active_logl=active_logl, loglstar=float(active_logl.min()))
    out.update(sd_arrays(t.netG))
    np.savez_compressed(os.path.join(HERE, 'bench_%s.npz' % tag), **out)
    z, ld = t.forward(active_u.astype(np.float32), to_numpy=True)
    print(tag, 'loglstar', out['loglstar'], 'latent std', z.std(), 'best val loss', t.best_validation_loss)


def mh(tag, d, rho):
    np.random.seed(0)
    torch.manual_seed(0)
    cov = (1 - rho) * np.eye(d) + rho * np.ones((d, d))
    samples = np.random.multivariate_normal(np.zeros(d), cov, size=4000)
    mean, std = samples.mean(0), samples.std(0)
    t = Trainer(d, hidden_dim=16, num_layers=1, num_blocks=3, flow='nvp', learning_rate=0.001, log_dir=None,
                log_level=logging.WARNING)
    t.train((samples - mean) / std, max_iters=EPOCHS, jitter=0.01)
    out = dict(d=d, rho=rho, epochs=EPOCHS, mean=mean, std=std)
    out.update(sd_arrays(t.netG))
    np.savez_compressed(os.path.join(HERE, 'bench_%s.npz' % tag), **out)
    print(tag, 'best val loss', t.best_validation_loss)


if __name__ == '__main__':
    hard('c2', rl.Himmelblau(2), olike.Himmelblau(2), 2, 5.0)
    hard('c3', rl.GaussianMix(10), olike.GaussianMix(10), 10, 10.0)
    hard('c4', rl.Rosenbrock(30), olike.Rosenbrock(30), 30, 5.0)
    mh('c5', 50, 0.99)
