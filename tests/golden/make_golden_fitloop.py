"""Fit-loop golden: the decisions of the REAL reference's Trainer.train (nnest/trainer.py:170-244) on scripted validation
losses -- best epoch, the epoch at which patience ran out, and WHICH epoch's weights it finally loads.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_fitloop.py
Trainer._train adds 1 to every weight (so the weights count the epochs), Trainer._validate returns the scripted loss;
everything else is the reference's own loop.  Writes fit_loop.npz.
"""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle.refload import load_reference  # noqa: E402

load_reference()
from nnest.trainer import Trainer  # noqa: E402

rng = np.random.RandomState(3)
CASES = {
    'improving': (list(np.linspace(5, 1, 12)), 12, 50),
    'patience': ([5, 4, 3] + [3.5] * 30, 30, 4),
    'patience_short': ([5, 6, 7, 8], 10, 1),
    'plateau': (list(rng.uniform(1, 2, size=40)), 40, 6),
    'one_epoch': ([2.0], 1, 50),
    'ties': ([3.0, 3.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0], 8, 3),
    'patience_zero': ([5.0, 4.0, 3.0], 3, 0),
}
out = {}
for name, (vals, max_iters, patience) in CASES.items():
    torch.manual_seed(0)
    np.random.seed(0)
    t = Trainer(3, flow='nvp', log_dir=None, log_level=logging.ERROR, batch_size=10)
    w0 = torch.nn.utils.parameters_to_vector(t.netG.parameters()).detach().clone()
    state = {'epoch': 0}

    def fake_train(epoch, loader, jitter=0.0, l2_norm=0.0, t=t, state=state):
        with torch.no_grad():
            for p in t.netG.parameters():
                p.add_(1.0)
        state['epoch'] = epoch
        return 1.0

    def fake_validate(epoch, loader, vals=vals):
        return float(vals[min(epoch, len(vals)) - 1])

    t._train, t._validate = fake_train, fake_validate
    t.train(np.random.uniform(-1, 1, size=(50, 3)), max_iters=max_iters, jitter=0.01, patience=patience)
    w = torch.nn.utils.parameters_to_vector(t.netG.parameters()).detach()
    kept = float((w - w0).mean())
    out[name + '/vals'] = np.array(vals, dtype=np.float64)
    out[name + '/max_iters'], out[name + '/patience'] = max_iters, patience
    out[name + '/best_epoch'] = t.best_validation_epoch
    out[name + '/epochs_run'] = state['epoch']
    out[name + '/kept_epoch'] = int(round(kept))
    out[name + '/total_iters'] = t.total_iters
    print(name, t.best_validation_epoch, state['epoch'], kept, t.total_iters)
np.savez_compressed(os.path.join(HERE, 'fit_loop.npz'), **out)
