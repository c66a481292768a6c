"""CPU: the flow-fitting oracle (oracle/train.py: hand-derived backward pass + Adam) against goldens recorded from the
real reference's Trainer._train / _validate (tests/golden/make_golden_train.py)."""
import numpy as np
import pytest

from oracle import train as otrain
from helpers import load

CASES = ['d2', 'd5_jit', 'd30', 'd10_big', 'd7_h32_l2_b2', 'd50']


def arch(g):
    return int(g['d']), int(g['hidden']), int(g['layers']), int(g['blocks'])


def flat_of(g, prefix):
    sd = {k[len(prefix) + 1:]: g[k] for k in g.files if k.startswith(prefix + '/')}
    return otrain.flatten_state_dict(sd, int(g['blocks'])).astype(np.float64)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('name', CASES)
def test_gradient_matches_reference_autograd(name):
    g = load('train_%s.npz' % name)
    d, H, L, B = arch(g)
    w0 = flat_of(g, 'sd')
    batch = int(g['batch'])
    nll, grad = otrain.nll_and_grad(w0, g['x_train'][:batch], d, H, L, B)
    assert abs(nll.mean() - float(g['first_loss'])) <= 1e-5 * abs(float(g['first_loss']))
    assert rel(grad, flat_of(g, 'grad')) < 1e-5


@pytest.mark.parametrize('name', CASES)
def test_two_epochs_match_reference_trainer(name):
    g = load('train_%s.npz' % name)
    d, H, L, B = arch(g)
    w = flat_of(g, 'sd')
    opt = otrain.Adam(w.size, lr=float(g['lr']), betas=tuple(g['betas']), eps=float(g['eps']),
                      weight_decay=float(g['weight_decay']))
    jit = float(g['jitter'])
    w, tl = otrain.train_epoch(w, opt, g['x_train'], int(g['batch']), d, H, L, B, jit, g['noise'] if jit else None)
    assert abs(tl - float(g['train_loss'])) <= 1e-5 * abs(float(g['train_loss']))
    # Adam's first steps move every weight by ~lr whatever the gradient's size, so compare on the scale of the update
    w0 = flat_of(g, 'sd')
    ref = flat_of(g, 'after')
    assert np.abs(w - ref).max() < 2e-3 * np.abs(ref - w0).max()
    assert abs(otrain.validate(w, g['x_valid'], d, H, L, B) - float(g['val_loss'])) <= 2e-5 * abs(float(g['val_loss']))
    w, tl2 = otrain.train_epoch(w, opt, g['x_train'], int(g['batch']), d, H, L, B, jit, g['noise2'] if jit else None)
    assert abs(tl2 - float(g['train_loss2'])) <= 2e-5 * abs(float(g['train_loss2']))
    ref2 = flat_of(g, 'after2')
    assert np.abs(w - ref2).max() < 2e-3 * np.abs(ref2 - w0).max()
    assert abs(otrain.validate(w, g['x_valid'], d, H, L, B) - float(g['val_loss2'])) <= 5e-5 * abs(float(g['val_loss2']))
