"""GPU: the fused flow-fitting kernel (nnb_train_epoch, csrc/nnb_train.cuh) through the C ABI against goldens recorded
from the real reference's Trainer._train / _validate (tests/golden/make_golden_train.py) and the CPU oracle
(oracle/train.py).  Tolerances: gradients and losses 1e-5 relative (max-norm); weights after k Adam steps are compared
on the scale of the accumulated update (Adam's first steps move every weight by ~lr whatever the gradient's size)."""
import logging

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import train as otrain
from helpers import load

pytestmark = pytest.mark.gpu

CASES = ['d2', 'd5_jit', 'd30', 'd10_big', 'd7_h32_l2_b2', 'd50']


@pytest.fixture(scope='module')
def engine():
    from nnest_b200.engine import Engine
    return Engine(0)


def arch(g):
    return int(g['d']), int(g['hidden']), int(g['layers']), int(g['blocks'])


def flat_of(g, prefix):
    sd = {k[len(prefix) + 1:]: g[k] for k in g.files if k.startswith(prefix + '/')}
    return otrain.flatten_state_dict(sd, int(g['blocks']))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to('cuda', dtype=dtype).contiguous()


def opt_kw(g):
    return dict(lr=float(g['lr']), betas=tuple(float(b) for b in g['betas']), eps=float(g['eps']),
                weight_decay=float(g['weight_decay']))


@pytest.mark.parametrize('name', CASES)
def test_gradient_and_losses_match_reference_autograd(engine, name):
    g = load('train_%s.npz' % name)
    a = arch(g)
    assert engine.train_supported(*a)
    batch = int(g['batch'])
    w = dev(flat_of(g, 'sd'))
    m, v, gout = torch.zeros_like(w), torch.zeros_like(w), torch.zeros_like(w)
    tl, vs, grid = engine.train_epoch(a, w, m, v, 0, dev(g['x_train'][:batch]), dev(g['x_valid']), batch,
                                      grad_out=gout, **opt_kw(g))
    assert grid == (batch + 127) // 128
    assert abs(tl - float(g['first_loss'])) <= 1e-5 * abs(float(g['first_loss']))
    assert rel(gout.cpu().numpy(), flat_of(g, 'grad')) < 1e-5
    # Adam moments after one step from zero: m = (1 - beta1) (g + wd w0), v = (1 - beta2) (g + wd w0)^2
    g_tot = flat_of(g, 'grad').astype(np.float64) + float(g['weight_decay']) * flat_of(g, 'sd')
    assert rel(m.cpu().numpy(), 0.1 * g_tot) < 1e-5
    assert rel(v.cpu().numpy(), 0.001 * g_tot ** 2) < 2e-5
    # validation only: parameters untouched, value = sum of -log p of the oracle
    w2 = dev(flat_of(g, 'sd'))
    _, vs0, _ = engine.train_epoch(a, w2, None, None, 0, None, dev(g['x_valid']), batch, do_train=False)
    nll, _ = otrain.nll_and_grad(flat_of(g, 'sd'), g['x_valid'], *a, want_grad=False)
    assert abs(vs0 - nll.sum()) <= 1e-5 * abs(nll.sum())
    assert torch.equal(w2, dev(flat_of(g, 'sd')))


@pytest.mark.parametrize('name', CASES)
def test_two_epochs_match_reference_trainer(engine, name):
    g = load('train_%s.npz' % name)
    a = arch(g)
    batch, jit = int(g['batch']), float(g['jitter'])
    n, nv = g['x_train'].shape[0], g['x_valid'].shape[0]
    w0 = flat_of(g, 'sd')
    w = dev(w0)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    xt, xv = dev(g['x_train']), dev(g['x_valid'])
    steps = 0
    for ep, (after, tlk, vlk, nk) in enumerate([('after', 'train_loss', 'val_loss', 'noise'),
                                                ('after2', 'train_loss2', 'val_loss2', 'noise2')]):
        tl, vs, _ = engine.train_epoch(a, w, m, v, steps, xt, xv, batch, jitter=jit,
                                       noise=dev(g[nk]) if jit else None, **opt_kw(g))
        steps += (n + batch - 1) // batch
        assert abs(tl / n - float(g[tlk])) <= 2e-5 * abs(float(g[tlk]))
        ref = flat_of(g, after)
        assert np.abs(w.cpu().numpy() - ref).max() < 2e-3 * np.abs(ref - w0).max()
        assert abs(vs / nv / nv - float(g[vlk])) <= 5e-5 * abs(float(g[vlk]))


def test_identity_order_equals_explicit_permutation_and_reverse_differs(engine):
    g = load('train_d5_jit.npz')
    a = arch(g)
    n = g['x_train'].shape[0]
    outs = []
    for perm in (None, torch.arange(n, device='cuda'), torch.arange(n - 1, -1, -1, device='cuda')):
        w = dev(flat_of(g, 'sd'))
        m, v = torch.zeros_like(w), torch.zeros_like(w)
        engine.train_epoch(a, w, m, v, 0, dev(g['x_train']), dev(g['x_valid']), 100, perm=perm, **opt_kw(g))
        outs.append(w.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


def test_philox_jitter_is_deterministic_and_has_unit_variance(engine):
    """jitter * N(0, I) from the library's Philox stream: same (seed, epoch) -> same result, another epoch -> another
    draw; E[-log p] under the noise agrees with the oracle fed numpy normals (4096 x 5 draws)."""
    g = load('train_d5_jit.npz')
    a = arch(g)
    rng = np.random.RandomState(0)
    x = np.repeat(g['x_train'][:64], 64, axis=0)
    losses = []
    for seed, epoch in ((7, 1), (7, 1), (7, 2)):
        w = dev(flat_of(g, 'sd'))
        m, v = torch.zeros_like(w), torch.zeros_like(w)
        tl, _, _ = engine.train_epoch(a, w, m, v, 0, dev(x), None, x.shape[0], jitter=0.3, seed=seed, epoch=epoch,
                                      **opt_kw(g))
        losses.append(tl)
    assert losses[0] == losses[1] and losses[0] != losses[2]
    ref = [otrain.nll_and_grad(flat_of(g, 'sd'), x + 0.3 * rng.normal(size=x.shape), *a, want_grad=False)[0]
           for _ in range(4)]
    mu, sd = np.mean([r.mean() for r in ref]), np.std(np.concatenate(ref)) / np.sqrt(x.shape[0])
    assert abs(losses[0] - mu) < 6 * sd and abs(losses[2] - mu) < 6 * sd


def test_many_ctas_per_minibatch_match_oracle(engine):
    """batch 4096 on 32 CTAs (cooperative launch, gradient exchange through global memory) against the float64 oracle."""
    g = load('train_d10_big.npz')
    a = arch(g)
    rng = np.random.RandomState(3)
    x = np.concatenate([g['x_train']] * 9)[:8192 + 1000] + 0.01 * rng.normal(size=(9000, 10)).astype(np.float32)
    x = x[:8192 + 808].astype(np.float32)
    w0 = flat_of(g, 'sd')
    w = dev(w0)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    tl, vs, grid = engine.train_epoch(a, w, m, v, 0, dev(x), dev(g['x_valid']), 4096, **opt_kw(g))
    assert grid == 32
    opt = otrain.Adam(w0.size, lr=float(g['lr']), betas=tuple(g['betas']), eps=float(g['eps']),
                      weight_decay=float(g['weight_decay']))
    wr, tlr = otrain.train_epoch(w0.astype(np.float64), opt, x, 4096, *a)
    assert abs(tl / x.shape[0] - tlr) <= 2e-5 * abs(tlr)
    assert np.abs(w.cpu().numpy() - wr).max() < 2e-3 * np.abs(wr - w0).max()
    assert rel(m.cpu().numpy(), opt.m) < 1e-4
    assert abs(vs / 200 / 200 - otrain.validate(wr, g['x_valid'], *a)) <= 5e-5 * abs(otrain.validate(wr, g['x_valid'], *a))


def test_unsupported_architecture_is_an_error(engine):
    assert not engine.train_supported(5, 64, 1, 3)
    from nnest_b200 import _lib as L
    w = torch.zeros(otrain.net_floats(5, 64, 1) * 6, device='cuda')
    with pytest.raises(L.NNBError):
        engine.train_epoch((5, 64, 1, 3), w, w.clone(), w.clone(), 0, torch.zeros((10, 5), device='cuda'), None, 10)


def test_mean_nn_distance_matches_kdtree(engine):
    import scipy.spatial
    rng = np.random.RandomState(0)
    for n, d in ((1000, 2), (777, 30), (300, 50)):
        x = rng.uniform(-1, 1, size=(n, d))
        x[5] = x[6]                                  # duplicates give distance 0, as the k-d tree does
        dists, _ = scipy.spatial.cKDTree(x).query(x, 2)
        got = engine.mean_nn_distance(dev(x, torch.float64))
        assert abs(got - dists[:, 1].mean()) <= 1e-12 * dists[:, 1].mean()
        assert abs(0.5 * got - np.mean(dists)) <= 1e-12


def test_trainer_uses_fused_kernel_and_fits(tmp_path):
    """Trainer.train (trainer.py:134-245) on the fused path: the loss falls, the sampling kernels see the new weights,
    netG.state_dict() holds them, and the autograd fallback reaches a comparable fit."""
    from nnest_b200 import Trainer
    np.random.seed(0)
    torch.manual_seed(0)
    cov = np.array([[1.0, 0.8], [0.8, 1.0]])
    x = np.random.multivariate_normal([1.0, -1.0], cov, size=2000)
    t = Trainer(2, flow='nvp', log_dir=str(tmp_path), log_level=logging.WARNING, learning_rate=0.001)
    assert t._fused
    before = -t.log_probs(x.astype(np.float32)).mean().item()
    launches0 = t.engine.gpu_launches
    t.train(x, max_iters=60, jitter=-1)
    assert t.engine.gpu_launches - launches0 >= 60
    after = -t.log_probs(x.astype(np.float32)).mean().item()
    assert after < before - 0.5 and after < 2.6          # entropy of the target: 2.33
    z, _ = t.forward(x.astype(np.float32))
    zt, _ = t.netG.forward(torch.from_numpy(x.astype(np.float32)).cuda())
    assert np.abs(z.cpu().numpy() - zt.detach().cpu().numpy()).max() < 1e-4
    assert t.best_validation_epoch >= 1


def test_ragged_and_degenerate_sizes(engine):
    """batch larger than the dataset, a dataset smaller than one CTA tile, no validation set, an empty epoch."""
    g = load('train_d2.npz')
    a = arch(g)
    w0 = flat_of(g, 'sd')
    kw = opt_kw(g)
    # (i) batch_size > n_train: one Adam step on all 37 samples == the oracle's
    x = g['x_train'][:37]
    w = dev(w0)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    tl, vs, grid = engine.train_epoch(a, w, m, v, 0, dev(x), None, 100, **kw)
    assert grid == 1 and vs == 0.0
    opt = otrain.Adam(w0.size, lr=kw['lr'], betas=kw['betas'], eps=kw['eps'], weight_decay=kw['weight_decay'])
    wr, tlr = otrain.train_epoch(w0.astype(np.float64), opt, x, 100, *a)
    assert abs(tl / 37 - tlr) <= 2e-5 * abs(tlr)
    assert np.abs(w.cpu().numpy() - wr).max() < 2e-3 * np.abs(wr - w0).max()
    # (ii) batch_size 1: 5 steps
    w = dev(w0)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    engine.train_epoch(a, w, m, v, 0, dev(x[:5]), dev(g['x_valid']), 1, **kw)
    opt = otrain.Adam(w0.size, lr=kw['lr'], betas=kw['betas'], eps=kw['eps'], weight_decay=kw['weight_decay'])
    wr, _ = otrain.train_epoch(w0.astype(np.float64), opt, x[:5], 1, *a)
    assert np.abs(w.cpu().numpy() - wr).max() < 2e-3 * np.abs(wr - w0).max()
    # (iii) empty training set: nothing changes, validation still evaluated
    w = dev(w0)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    tl, vs, _ = engine.train_epoch(a, w, m, v, 0, None, dev(g['x_valid']), 100, **kw)
    assert tl == 0.0 and vs > 0.0 and np.array_equal(w.cpu().numpy(), w0) and float(m.abs().max()) == 0.0


def test_queueing_epochs_ahead_does_not_change_the_fit(tmp_path):
    """Trainer.train queues epoch e + 1 before it reads the losses of epoch e (nnb_train_epoch_begin / _end).  With the same
    seeds the weights, the best epoch and the losses are bit-identical to the one-epoch-at-a-time loop -- also when patience
    runs out and when the fit is called again."""
    from nnest_b200 import Trainer
    x = np.random.RandomState(4).normal(size=(3000, 4)) * np.array([1.0, 0.5, 2.0, 0.1])
    out = []
    for lookahead in (True, False):
        np.random.seed(7)
        torch.manual_seed(7)
        t = Trainer(4, flow='nvp', log_dir=str(tmp_path / str(lookahead)), log_level=logging.ERROR, batch_size=256,
                    learning_rate=0.01)
        t._lookahead = lookahead
        t.train(x, max_iters=40, jitter=-1, patience=3)
        first = (t.best_validation_epoch, t.best_validation_loss, t.total_iters)
        t.train(x[:2000], max_iters=12, jitter=0.02)
        w = torch.nn.utils.parameters_to_vector(list(t.netG.parameters())).detach().cpu().numpy()
        out.append((first, t.best_validation_epoch, t.best_validation_loss, t.total_iters, t._adam_step, w))
    a, b = out
    assert a[:5] == b[:5], (a[:5], b[:5])
    assert np.array_equal(a[5], b[5])


def test_trainer_tiny_dataset_and_large_batch(tmp_path):
    """Trainer.train with fewer samples than batch_size and with a validation split of one sample."""
    from nnest_b200 import Trainer
    np.random.seed(1)
    torch.manual_seed(1)
    x = np.random.normal(size=(9, 3))
    t = Trainer(3, flow='nvp', log_dir=str(tmp_path), log_level=logging.WARNING, batch_size=100)
    t.train(x, max_iters=5, jitter=0.0)
    assert np.isfinite(t.best_validation_loss)
    z, ld = t.forward(x.astype(np.float32))
    assert torch.isfinite(z).all() and torch.isfinite(ld).all()


def test_trainer_recovers_from_a_diverged_epoch(tmp_path):
    """A non-finite loss (here: provoked by a NaN in the training data of one call) must not poison later fits: the
    best weights are restored and the Adam moments restarted."""
    from nnest_b200 import Trainer
    np.random.seed(2)
    torch.manual_seed(2)
    x = np.random.normal(size=(600, 3))
    t = Trainer(3, flow='nvp', log_dir=str(tmp_path), log_level=logging.ERROR, batch_size=100)
    t.train(x, max_iters=5, jitter=0.0)
    good = t.best_validation_loss
    bad = x.copy()
    bad[5, 1] = np.nan
    t.train(bad, max_iters=3, jitter=0.0)
    assert float(t._adam_m.abs().max()) == 0.0 or torch.isfinite(t._adam_m).all()
    z, ld = t.forward(x.astype(np.float32))
    assert torch.isfinite(z).all() and torch.isfinite(ld).all()          # weights are still the last good ones
    t.train(x, max_iters=5, jitter=0.0)
    assert np.isfinite(t.best_validation_loss) and t.best_validation_loss <= good * 1.05 + 1e-3


def test_gradient_only_shares_sum_to_the_minibatch_gradient(engine):
    """Data-parallel mode (nnb_train_args.grad_only): the gradients of two ranks' shares of a mini-batch, each scaled by
    1 / batch_total, add up to the gradient of the whole mini-batch; parameters and Adam moments stay untouched."""
    g = load('train_d10_big.npz')
    a = arch(g)
    x = dev(g['x_train'][:1000])
    w = dev(flat_of(g, 'sd'))
    w0 = w.clone()
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    full = torch.zeros_like(w)
    tl, _, _ = engine.train_epoch(a, w.clone(), m, v, 0, x, None, 1000, grad_out=full, **opt_kw(g))
    parts, loss = [], 0.0
    for lo, hi in ((0, 430), (430, 1000)):
        gp = torch.zeros_like(w)
        tlp, _, _ = engine.train_epoch(a, w, None, None, 0, x[lo:hi], None, hi - lo, grad_out=gp, grad_only=True,
                                       batch_total=1000, **opt_kw(g))
        parts.append(gp)
        loss += tlp
    assert torch.equal(w, w0)
    scale = full.abs().max().item()
    assert (parts[0] + parts[1] - full).abs().max().item() < 2e-6 * scale
    assert abs(loss - tl) < 1e-5 * abs(tl)


@pytest.mark.parametrize('n,d', [(3000, 2), (5000, 10), (4097, 30), (900, 17)])
def test_mean_nn_distance_prefilter_is_exact(engine, n, d):
    """float32 prefilter + float64 refinement == float64 brute force (numpy), also for clustered points far from the origin,
    duplicated rows (distance 0) and near-ties."""
    rng = np.random.RandomState(n + d)
    x = rng.normal(size=(n, d)) * 1e-3 + 0.7                  # a tight cloud away from the origin
    x[5] = x[6]                                               # an exact duplicate
    x[7] = x[8] + 1e-9                                        # a near-duplicate below float32 resolution
    got = engine.mean_nn_distance(torch.from_numpy(x).cuda())
    best = np.full(n, np.inf)
    for lo in range(0, n, 512):
        dd = ((x[lo:lo + 512, None, :] - x[None, :, :]) ** 2).sum(-1)
        dd[np.arange(min(512, n - lo)), np.arange(lo, min(n, lo + 512))] = np.inf
        best[lo:lo + 512] = dd.min(1)
    want = np.sqrt(best).mean()
    assert abs(got - want) <= 1e-12 * want


@pytest.mark.parametrize('n,d,kind', [(20000, 30, 'uniform'), (9000, 12, 'two_clusters'), (8300, 52, 'uniform'),
                                      (8200, 1, 'uniform'), (9000, 30, 'tight'), (8300, 36, 'uniform')])
def test_mean_nn_distance_tensor_core_path_is_exact(engine, n, d, kind):
    """n >= 8192: the pairs are rated on the tensor cores (3xTF32, csrc/nnb_nn_tc.cuh) and the candidates within the error
    bound of the running minimum re-evaluated in float64 -- the same minimum as the float64 brute force (numpy), for a
    typical live set, for two far-apart tight clusters (the slack then covers a whole cluster: many exact evaluations, same
    answer), for K = 56 (one query tile per CTA), K = 40 (two) and K = 8 operands, with duplicates and near-ties."""
    rng = np.random.RandomState(n + d)
    if kind == 'uniform':
        x = rng.uniform(-1, 1, size=(n, d))
    elif kind == 'tight':
        x = rng.normal(size=(n, d)) * 1e-3 + 0.7
    else:
        x = rng.normal(size=(n, d)) * 1e-3
        x[n // 2:] += 10.0
    x[5] = x[6]
    x[7] = x[8] + 1e-9
    got = engine.mean_nn_distance(torch.from_numpy(x).cuda())
    best = np.full(n, np.inf)
    for lo in range(0, n, 256):
        dd = ((x[lo:lo + 256, None, :] - x[None, :, :]) ** 2).sum(-1)
        dd[np.arange(min(256, n - lo)), np.arange(lo, min(n, lo + 256))] = np.inf
        best[lo:lo + 256] = dd.min(1)
    want = np.sqrt(best).mean()
    assert abs(got - want) <= 1e-12 * want
