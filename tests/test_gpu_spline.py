"""GPU: the neural-spline flow (the reference's default flow='spline') through the C ABI:
flow maps against goldens recorded from the real reference, the fused MCMC step against the oracle on dumped noise, the
Trainer facade / fitting, and the reference's own integration test (tests/test_nested.py) as shipped."""
import logging

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import likelihoods as olike
from oracle import mcmc as omcmc
from oracle import spline as ospline
from helpers import load, rel_err

pytestmark = pytest.mark.gpu

CASES = ['d2', 'd3', 'd5', 'd10', 'd4_h8_b2']
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope='module')
def engine():
    from nnest_b200.engine import Engine
    return Engine(0)


def _install(engine, g):
    w = ospline.SplineWeights.from_golden(g)
    d, hidden, blocks = int(g['d']), int(g['hidden']), int(g['blocks'])
    engine.set_flow_spline(ospline.pack_for_kernel(w, hidden), d, hidden, blocks, w.K, w.B)
    return w


@pytest.mark.parametrize('name', CASES)
def test_spline_flow_kernels_match_reference(engine, name):
    g = load('spline_%s.npz' % name)
    _install(engine, g)
    z, ld = engine.flow_forward(dev(g['x']))
    assert rel_err(z.cpu().numpy(), g['fwd_z']) < 1e-5
    assert np.abs(ld.cpu().numpy() - g['fwd_ld']).max() < 2e-5 * max(1.0, np.abs(g['fwd_ld']).max())
    x, ldx = engine.flow_inverse(dev(g['zin']))
    assert rel_err(x.cpu().numpy(), g['inv_x']) < 1e-5
    assert np.abs(ldx.cpu().numpy() - g['inv_ld']).max() < 2e-5 * max(1.0, np.abs(g['inv_ld']).max())
    # tests/test_flows.py:56-72: round trip and log-det antisymmetry
    xr, ldr = engine.flow_inverse(z)
    assert np.abs(xr.cpu().numpy() - g['x']).max() < 2e-5 and (ld + ldr).abs().max().item() < 5e-5
    # strided (chain-minor) views
    zt, _ = engine.flow_forward(dev(np.ascontiguousarray(g['x'].T)).t())
    assert torch.equal(zt, z)


def test_spline_empty_half_flags(engine):
    g = load('spline_d2.npz')
    _install(engine, g)
    z = dev(np.array([[0.1, 0.2], [5.0, 0.3], [7.0, -9.0], [0.0, 3.5]], dtype=np.float32))
    flags = engine.flow_empty_halves(z, inverse=True).cpu().numpy()
    assert flags[0] == 0 and flags[2] == 1           # everything outside the tail bound: the reference raises on it alone
    x, ld = engine.flow_inverse(z)
    assert torch.isfinite(x).all() and torch.isfinite(ld).all()


@pytest.mark.parametrize('name,mode', [('d5', 0), ('d10', 0), ('d3', 1)])
def test_spline_mcmc_matches_oracle_on_dumped_noise(engine, name, mode):
    """Fused MCMC step with the spline flow: free-running Philox noise dumped by the kernel, replayed by the oracle
    (oracle/mcmc.py with oracle/spline.py as the flow)."""
    g = load('spline_%s.npz' % name)
    w = _install(engine, g)
    d, n, steps = int(g['d']), 600, 8
    like = olike.Rosenbrock(d)
    rng = np.random.RandomState(d)
    if mode == 0:
        engine.set_target(d, 0, [], t_scale=5.0, t_shift=0.0, prior_kind=1, prior_lo=-1.0, prior_hi=1.0)
        # start inside the flow's image of a modest latent ball so that the chains move
        z0 = (0.5 * rng.normal(size=(n, d))).astype(np.float32)
        u0, _ = ospline.flow_inverse(w, z0)
        u0 = np.clip(u0, -0.9, 0.9).astype(np.float64)
        logl0 = like.batch(5 * u0)
        loglstar = float(np.percentile(logl0, 25))
        st, _, _ = engine.mcmc_init(n, init_u=dev(u0.astype(np.float32).T), init_logl=dev(logl0), seed=3)
        kw = dict(init_samples=u0, init_loglikes=logl0, loglstar=loglstar, step_size=0.3, dynamic_step_size=True)
        target = omcmc.Target(like, transform=lambda x: 5 * x, prior=olike.UniformPrior(d, -1, 1), transform_prior=False)
        out = engine.mcmc_run(st, steps, mode=0, loglstar=loglstar, step_size=0.3, dynamic_step_size=True, seed=3,
                              trace=True, dump_noise=True)
    else:
        engine.set_target(d, 0, [], t_scale=2.0, t_shift=0.0, compute_f64=True, prior_kind=2, prior_lo=-5.0, prior_hi=5.0)
        z0 = (0.5 * rng.normal(size=(n, d))).astype(np.float32)
        st, _, _ = engine.mcmc_init(n, init_z=dev(z0.T), seed=3)
        kw = dict(init_z=z0, loglstar=None, step_size=0.4)
        target = omcmc.Target(like, transform=lambda x: 2.0 * x.astype(np.float64), prior=olike.UniformPrior(d, -5, 5),
                              transform_prior=True)
        out = engine.mcmc_run(st, steps, mode=1, step_size=0.4, seed=3, trace=True, dump_noise=True)
    ref = omcmc.mcmc_sample(w, target, steps, omcmc.ReplayNoise(out['normals'].cpu().numpy(),
                                                                 out['uniforms'].cpu().numpy()), flow=ospline, **kw)
    latent = out['trace_z'].permute(2, 0, 1).cpu().numpy()
    samples = out['trace_x'].permute(2, 0, 1).cpu().numpy()
    moved_ref = np.any(ref[1][:, 1:] != ref[1][:, :-1], axis=2)
    moved = np.any(latent[:, 1:] != latent[:, :-1], axis=2)
    same = np.all(moved == moved_ref, axis=1)
    assert (~same).sum() <= 3                       # near-ties only
    assert moved.sum() > 0.05 * moved.size          # the test moves chains
    assert rel_err(latent[same], ref[1][same]) < 1e-5
    assert rel_err(samples[same], ref[0][same]) < 1e-5
    assert torch.equal(st.z, out['trace_z'][-1]) and torch.equal(st.x, out['trace_x'][-1])


def test_spline_trainer_facade_and_fit(tmp_path):
    from nnest_b200 import Trainer
    g = load('spline_d5.npz')
    t = Trainer(5, flow='spline', hidden_dim=int(g['hidden']), num_blocks=int(g['blocks']), log_dir=str(tmp_path),
                log_level=logging.WARNING)
    sd = {k[3:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith('sd/')}
    t.load_state_dict(sd, permutations=[g['P/%d' % k] for k in range(int(g['blocks']))])
    z, ld = t.forward(g['x'])
    assert z.is_cuda and rel_err(z.cpu().numpy(), g['fwd_z']) < 1e-5
    xn, ldn = t.inverse(g['zin'], to_numpy=True)
    assert rel_err(xn, g['inv_x']) < 1e-5
    # kernels and the autograd module agree
    zt, ldt = t.netG.forward(torch.from_numpy(g['x']).cuda())
    assert rel_err(zt.detach().cpu().numpy(), g['fwd_z']) < 1e-5
    with pytest.raises(ValueError):                   # RQS on an empty selection (networks.py:464-465)
        t.inverse(np.full((1, 5), 50.0, dtype=np.float32))
    # fitting: data-dependent ActNorm initialisation + Adam through the CUDA-graph step
    torch.manual_seed(0)
    np.random.seed(0)
    t2 = Trainer(2, flow='spline', hidden_dim=16, num_blocks=3, log_dir=str(tmp_path / 'fit'), learning_rate=0.001,
                 log_level=logging.WARNING)
    x = np.random.normal(size=(800, 2)) * np.array([0.2, 0.05]) + np.array([0.3, -0.2])
    t2.train(x, max_iters=40, jitter=-1.0)
    after = -t2.log_probs(x.astype(np.float32)).mean().item()
    assert after < -1.0                                # entropy of the target Gaussian is -1.77 nats
    with torch.no_grad():
        ref = t2.netG.log_probs(torch.from_numpy(x.astype(np.float32)).cuda()).cpu().numpy()
    assert np.allclose(t2.log_probs(x.astype(np.float32), to_numpy=True), ref, rtol=1e-4, atol=1e-4)


def test_reference_test_nested_as_shipped(tmp_path):
    """/root/reference/tests/test_nested.py:10-19 verbatim (flow='spline', 1000 live points, 10 chains, fixed step size),
    with nnest_b200 in place of nnest."""
    from nnest_b200 import NestedSampler
    from nnest_b200.likelihoods import Rosenbrock
    np.random.seed(0)
    torch.manual_seed(0)
    max_evidence_error = 0.2
    transform = lambda x: 5 * x
    like = Rosenbrock(2)
    sampler = NestedSampler(2, like, transform=transform,
                            num_live_points=1000, hidden_dim=16,
                            num_layers=1, num_blocks=3, num_slow=0,
                            flow='spline', log_dir=str(tmp_path), log_level=logging.WARNING)
    sampler.run(mcmc_num_chains=10, mcmc_dynamic_step_size=False)
    diff = sampler.logz + 5.80
    assert np.abs(diff) <= max_evidence_error
