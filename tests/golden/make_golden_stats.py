"""Golden fixture for the chain diagnostics, recorded from the REAL reference (nnest/utils/evaluation.py:6-73, the
functions Sampler._chain_stats calls, nnest/sampler.py:474-492).  Build container only (needs /root/reference):
    python tests/golden/make_golden_stats.py
chain_stats.npz: chains (b, t, d) float32 with repeated points (rejected proposals), mean / std as _chain_stats derives
them, and the reference's acceptance_rate, effective_sample_size (divides by `std`, sic), mean_jump_distance."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle.refload import load_reference  # noqa: E402

load_reference()
from nnest.utils.evaluation import acceptance_rate, effective_sample_size, mean_jump_distance  # noqa: E402

rng = np.random.RandomState(4)
out = {}
for tag, (b, t, d, rho) in {'a': (6, 80, 3, 0.9), 'b': (40, 200, 5, 0.97), 'c': (130, 33, 2, 0.5)}.items():
    x = np.zeros((b, t, d), dtype=np.float32)
    x[:, 0] = rng.normal(size=(b, d))
    for s in range(1, t):
        prop = rho * x[:, s - 1] + np.sqrt(1 - rho ** 2) * rng.normal(size=(b, d)).astype(np.float32)
        keep = rng.uniform(size=b) < 0.4                      # rejected proposal: the point is repeated
        x[:, s] = np.where(keep[:, None], x[:, s - 1], prop)
    scale, shift = rng.uniform(0.5, 3.0, size=d), rng.normal(size=d)
    v = x * scale + shift                                      # float64, like samples * std + mean (mcmc.py:117)
    mean = np.mean(np.reshape(v, (-1, d)), axis=0)
    std = np.std(np.reshape(v, (-1, d)), axis=0)
    out.update({tag + '_x': x, tag + '_scale': scale, tag + '_shift': shift, tag + '_mean': mean, tag + '_std': std,
                tag + '_acceptance': acceptance_rate(v), tag + '_ess': effective_sample_size(v, mean, std),
                tag + '_jump': mean_jump_distance(v)})
    print(tag, out[tag + '_acceptance'], out[tag + '_ess'], out[tag + '_jump'])
np.savez_compressed(os.path.join(HERE, 'chain_stats.npz'), **out)
