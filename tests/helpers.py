"""Shared helpers for the test-suite (fixtures -> oracle objects)."""
import os

import numpy as np

from oracle import flow as oflow
from oracle import likelihoods as olike

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def state_dict_of(g):
    return {k[3:]: g[k] for k in g.files if k.startswith('sd/')}


def weights_of(g):
    scale = str(g['scale']) if 'scale' in g.files and g['scale'].dtype.kind in 'US' else ''
    return oflow.NVPWeights.from_state_dict(state_dict_of(g), int(g['d']), scale=scale)


def rel_err(a, b):
    """max |a-b| / max|b|  (scale-relative error; the 1e-5 criterion of north_star)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


LIKE_CASES = {
    'rosenbrock2': (lambda: olike.Rosenbrock(2), 5.0),
    'rosenbrock30': (lambda: olike.Rosenbrock(30), 5.0),
    'himmelblau': (lambda: olike.Himmelblau(2), 5.0),
    'gaussian10': (lambda: olike.Gaussian(10, 0.99, lim=3), 3.0),
    'gaussian50': (lambda: olike.Gaussian(50, 0.99, lim=3), 3.0),
    'eggbox': (lambda: olike.Eggbox(2), None),
    'mixture10': (lambda: olike.GaussianMix(10), 10.0),
    'mixture2': (lambda: olike.GaussianMix(2), 10.0),
    'shell5': (lambda: olike.GaussianShell(5), 5.0),
}
