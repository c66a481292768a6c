"""Chain-file golden: the bytes the REAL reference's Sampler._save_samples (nnest/sampler.py:494-527) writes for fixed inputs.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_chain.py
Writes chain_files.npz: the inputs (samples, loglikes, weights, derived, param names) and, as uint8 arrays, the text of the
files the reference wrote -- single chain with header and derived columns, single chain without weights, and the
multi-chain layout chain_<k>.txt.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle.refload import load_reference  # noqa: E402

load_reference()
from nnest.sampler import Sampler  # noqa: E402

rng = np.random.RandomState(20)
out = {}


def run(tag, smp, lgl, w, der, names):
    d = tempfile.mkdtemp(prefix='nnest_chain_')
    me = types.SimpleNamespace(param_names=names, logs={'chains': d})
    Sampler._save_samples(me, smp, lgl, weights=w, derived_samples=der)
    for f in sorted(os.listdir(d)):
        out['%s/%s' % (tag, f)] = np.frombuffer(open(os.path.join(d, f), 'rb').read(), dtype=np.uint8)
    out[tag + '/samples'], out[tag + '/loglikes'] = smp, lgl
    if w is not None:
        out[tag + '/weights'] = w
    if der is not None:
        out[tag + '/derived'] = der
    if names is not None:
        out[tag + '/names'] = np.array(names)


n, d = 400, 5
smp = rng.normal(size=(n, d)) * 10.0 ** rng.uniform(-6, 6, size=(n, 1))
smp[0, 0], smp[1, 1], smp[2, 2], smp[3, 3] = 0.0, 9.999995e4, 1.0000005, -9.9999949999e-7
lgl = rng.normal(size=n) * 100
lgl[5], lgl[6] = np.inf, -np.inf
w = rng.uniform(size=n) * 10.0 ** rng.uniform(-45, 0, size=n)      # some fall below min_weight = 1e-30
der = rng.normal(size=(n, 2))
run('named', smp, lgl, w, der, ['a', 'b', 'c', 'd', 'e'])
run('plain', smp[:50], lgl[:50], None, None, None)
run('multi', smp[:60].reshape(3, 20, d), lgl[:60].reshape(3, 20), w[:60].reshape(3, 20), None, None)
np.savez_compressed(os.path.join(HERE, 'chain_files.npz'), **out)
print({k: v.shape for k, v in out.items()})
