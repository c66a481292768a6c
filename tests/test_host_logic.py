"""CPU tests of the host-side logic that needs no GPU: transform probing, state_dict flattening, the run-directory
layout, error behaviour without a CUDA device."""
import logging
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import flow as oflow
from oracle import train as otrain
from helpers import load, state_dict_of, weights_of


def test_probe_affine_transform_recovers_the_reference_maps():
    """the maps of examples/nested/run.py:25-44 and nnest/mcmc.py:111"""
    from nnest_b200.sampler import probe_affine_transform
    for fn, sc, sh in [(lambda x: 5 * x, 5.0, 0.0), (lambda x: x * 5 * np.pi, 5 * np.pi, 0.0),
                       (lambda x: 10 * x, 10.0, 0.0)]:
        scale, shift, promotes = probe_affine_transform(fn, 4)
        assert np.allclose(scale, sc, rtol=1e-15) and np.allclose(shift, sh) and not promotes
    std, mean = np.array([0.5, 2.0, 3.0]), np.array([1.0, -1.0, 0.25])
    scale, shift, promotes = probe_affine_transform(lambda x: x * std + mean, 3)
    assert np.allclose(scale, std) and np.allclose(shift, mean)
    assert promotes                      # float64 arrays promote float32 input: the likelihood then runs in float64
    assert probe_affine_transform(None, 3) == (None, None, False)


@pytest.mark.parametrize('fn', [lambda x: x ** 2, lambda x: np.sin(x), lambda x: x @ np.array([[1.0, 0.5], [0.0, 1.0]]),
                                lambda x: x[:, :1]])
def test_probe_rejects_non_affine_or_mixing_transforms(fn):
    from nnest_b200.sampler import probe_affine_transform
    with pytest.raises(NotImplementedError):
        probe_affine_transform(fn, 2)


@pytest.mark.parametrize('name', ['flow_d5.npz', 'flow_d7_h32_l2_b5.npz', 'flow_d6_translate.npz', 'flow_d6_constant.npz'])
def test_flatten_state_dict_matches_the_oracle_layout(name):
    """engine.flatten_state_dict (the nnb_set_flow input) == the oracle's own flattening of the same golden weights."""
    from nnest_b200.engine import flatten_state_dict
    g = load(name)
    scale = str(g['scale']) if g['scale'].dtype.kind in 'US' else ''
    flat, d, hidden, nl, nb, flags = flatten_state_dict(state_dict_of(g), scale)
    w = weights_of(g)
    assert (d, hidden, nl, nb) == (w.d, w.hidden, w.num_layers, w.num_blocks)
    assert np.array_equal(flat, w.flat())
    if scale == '':
        assert np.array_equal(flat, otrain.flatten_state_dict(state_dict_of(g), nb).astype(np.float32))
        assert flat.size == 2 * nb * otrain.net_floats(d, hidden, nl)


def test_run_directory_layout(tmp_path):
    """<log_dir>/runN/{info,results,chains,checkpoint,plots} (nnest/utils/logger.py:38-75)"""
    from nnest_b200.utils.logger import get_or_create_run_dir
    a = get_or_create_run_dir(str(tmp_path / 'logs'))
    b = get_or_create_run_dir(str(tmp_path / 'logs'))
    assert a['created'] and b['created']
    assert a['run_dir'].endswith('run1') and b['run_dir'].endswith('run2')
    for sub in ('info', 'results', 'chains', 'checkpoint', 'plots'):
        assert os.path.isdir(a[sub])
    c = get_or_create_run_dir(a['run_dir'])              # an existing run directory is resumed, not renumbered
    assert not c['created'] and c['run_dir'] == a['run_dir']
    d = get_or_create_run_dir(str(tmp_path / 'flat'), append_run_num=False)
    assert d['created'] and d['run_dir'].endswith('flat')


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour WITHOUT a CUDA device')
def test_samplers_and_trainer_refuse_to_run_without_a_gpu(tmp_path):
    from nnest_b200 import NestedSampler, Trainer
    from nnest_b200.likelihoods import Rosenbrock
    with pytest.raises(Exception):
        Trainer(2, flow='nvp', log_dir=None, log_level=logging.ERROR)
    with pytest.raises(Exception):
        NestedSampler(2, Rosenbrock(2), flow='nvp', log_dir=str(tmp_path), log_level=logging.ERROR)
    with pytest.raises(NotImplementedError):
        Trainer(2, flow='choleksy', log_dir=None)          # only 'nvp' and 'spline' are implemented on the device
    with pytest.raises(NotImplementedError):
        Trainer(2, flow='spline', num_slow=1, log_dir=None)


@pytest.mark.parametrize('arch', [(30, 16, 3, 1), (5, 32, 5, 2), (2, 16, 1, 1)])
def test_parameter_order_is_the_set_flow_layout(arch):
    """Trainer._sync_device hands the flat parameter vector (registration order) to nnb_set_flow: it must be the layout
    flatten_state_dict builds key by key from the state_dict (networks.py:262-282)."""
    from nnest_b200.networks import SingleSpeedNVP
    from nnest_b200.engine import flatten_state_dict
    d, h, b, l = arch
    g = SingleSpeedNVP(d, h, b, l, scale='', device=torch.device('cpu'))
    flat = torch.nn.utils.parameters_to_vector([p.detach() for p in g.parameters()]).numpy()
    want, dd, hh, ll, bb, flags = flatten_state_dict(g.state_dict(), '')
    assert (dd, hh, ll, bb, flags) == (d, h, l, b, 0) and np.array_equal(flat, want)


class _ScriptedEngine(object):
    """Stands in for the device in the test of the fit loop's HOST logic: an epoch adds 1 to every weight when it is queued
    and reports the scripted validation loss when it is collected.  Records what was queued and how many epochs were in
    flight."""

    def __init__(self, val_losses):
        self.device = torch.device('cpu')
        self.val = list(val_losses)
        self.begun, self.pending, self.max_in_flight, self.dropped = [], [], 0, 0
        self.gpu_launches = 0

    def train_supported(self, *arch):
        return True

    def set_flow(self, flat, d, hidden, num_layers, num_blocks, flags=0):
        self.installed = np.array(flat, copy=True)

    def mean_nn_distance(self, x):
        return 0.0

    def train_epoch_begin(self, arch, flat, m, v, step0, x_train, x_valid, batch, perm=None, epoch=0, **kw):
        assert len(self.pending) < 2                                     # the library's limit
        assert sorted(perm.tolist()) == list(range(x_train.shape[0]))
        k = len(self.begun)
        self.begun.append((int(epoch), int(step0), float(flat[0])))
        flat += 1.0
        self.pending.append(self.val[k] if k < len(self.val) else self.val[-1])
        self.max_in_flight = max(self.max_in_flight, len(self.pending))

    def train_epoch_end(self):
        v = self.pending.pop(0)
        return 1.0, v, 1

    def train_epoch_drain(self):
        self.dropped += len(self.pending)
        del self.pending[:]


FIT_CASES = ['improving', 'patience', 'patience_short', 'plateau', 'one_epoch', 'ties', 'patience_zero']


@pytest.mark.parametrize('lookahead', [True, False])
@pytest.mark.parametrize('case', FIT_CASES)
def test_fit_loop_queues_epochs_ahead_without_changing_the_decisions(case, lookahead, monkeypatch):
    """Trainer.train queues epoch e + 1 before it reads the losses of epoch e whenever epoch e cannot end the fit.  The
    decisions -- best epoch, the epoch at which patience runs out, WHICH epoch's weights are kept -- are those the REAL
    reference's loop (nnest/trainer.py:170-244) took on the same scripted validation losses (tests/golden/fit_loop.npz,
    make_golden_fitloop.py); no epoch is left in flight or run in excess; Adam steps and epoch ids are numbered as in the
    sequential loop."""
    from nnest_b200.trainer import Trainer
    g = load('fit_loop.npz')
    vals = [float(v) for v in g[case + '/vals']]
    max_iters, patience = int(g[case + '/max_iters']), int(g[case + '/patience'])
    best_epoch, ran, kept = int(g[case + '/best_epoch']), int(g[case + '/epochs_run']), int(g[case + '/kept_epoch'])
    rng = np.random.RandomState(3)
    eng = _ScriptedEngine(vals)
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)     # the scripted engine stands in for the device
    t = Trainer(3, flow='nvp', log_dir=None, log_level=logging.ERROR, batch_size=10, engine=eng)
    assert t._fused
    t._lookahead = lookahead
    w0 = torch.nn.utils.parameters_to_vector(list(t.netG.parameters())).detach().clone()
    x = rng.uniform(-1, 1, size=(50, 3))
    t.train(x, max_iters=max_iters, jitter=0.01, patience=patience)
    assert t.best_validation_epoch == best_epoch
    assert len(eng.begun) == ran and not eng.pending and eng.dropped == 0       # nothing queued in excess
    if lookahead and ran > 2 and patience > 1:
        assert eng.max_in_flight == 2                                            # ... and epochs WERE queued ahead
    if not lookahead:
        assert eng.max_in_flight == 1
    n_train = 50 - 5
    steps = (n_train + 9) // 10
    assert [b[0] for b in eng.begun] == list(range(1, ran + 1))                  # epoch ids = total_iters
    assert [b[1] for b in eng.begun] == [k * steps for k in range(ran)]          # Adam step numbering
    assert np.allclose([b[2] for b in eng.begun], [float(w0[0]) + k for k in range(ran)], atol=1e-4)   # e starts from e - 1
    w = torch.nn.utils.parameters_to_vector(list(t.netG.parameters())).detach()
    assert torch.allclose(w, w0 + kept, atol=1e-4)                               # the weights the reference keeps
    assert np.allclose(eng.installed, w.numpy())                                 # ... and installed in the sampling kernels
    assert t.total_iters == ran == int(g[case + '/total_iters']) and t._adam_step == ran * steps
    # a second fit continues the numbering
    eng.val = [1.0, 0.5]
    k0 = len(eng.begun)
    t.train(x, max_iters=2, jitter=0.01)
    assert [b[0] for b in eng.begun[k0:]] == [ran + 1, ran + 2] and eng.begun[k0][1] == ran * steps


def test_fit_loop_recovers_from_a_non_finite_epoch_with_one_queued_behind_it(monkeypatch):
    from nnest_b200.trainer import Trainer
    vals = [3.0, 2.0, float('nan'), 9.0, 1.5, 1.0]       # epoch 3 diverges while epoch 4 is already queued
    eng = _ScriptedEngine(vals)
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    t = Trainer(3, flow='nvp', log_dir=None, log_level=logging.ERROR, batch_size=10, engine=eng)
    w0 = torch.nn.utils.parameters_to_vector(list(t.netG.parameters())).detach().clone()
    t.train(np.random.RandomState(0).uniform(-1, 1, size=(50, 3)), max_iters=5, jitter=0.01)
    # queued: 1, 2, 3, 4 (dropped: it started from the diverged weights), then 4 again from the restored weights, 5
    assert [b[0] for b in eng.begun] == [1, 2, 3, 4, 4, 5]
    assert eng.begun[4][1] == 0 and abs(eng.begun[4][2] - float(w0[0]) - 2) < 1e-4    # Adam restarted, best weights restored
    assert t.best_validation_epoch == 5 and not eng.pending
    w = torch.nn.utils.parameters_to_vector(list(t.netG.parameters())).detach()
    assert torch.allclose(w, w0 + 2 + 2, atol=1e-4)                              # epochs 4 and 5 on top of the restored weights
