"""Golden fixtures for the neural-spline flow, recorded from the REAL reference
(nnest/networks.py:393-715: MLP, unconstrained_RQS / RQS, NSF_CL, Invertible1x1Conv, ActNorm, SingleSpeedSpline as built by
nnest/trainer.py:92-98).  Build container only (needs /root/reference):   python tests/golden/make_golden_spline.py

spline_<tag>.npz: netG.state_dict() AFTER the data-dependent ActNorm initialisation (the first forward call,
networks.py:689-695), the permutation matrices P of the 1x1 convolutions (plain attributes in the reference: not part of the
state_dict), inputs with points beyond the tail bound, and the reference's forward / inverse outputs and log-determinants.
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


def make_spline(tag, d, hidden=16, blocks=3, n=96, perturb=0.08, seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    t = Trainer(d, hidden_dim=hidden, num_blocks=blocks, flow='spline', log_dir=None, log_level=logging.WARNING)
    with torch.no_grad():
        for name, p in t.netG.named_parameters():
            if '.f1.' in name or '.f2.' in name:
                p.add_(perturb * torch.randn_like(p))
    init = (1.3 * np.random.normal(size=(256, d)) + 0.2).astype(np.float32)
    t.forward(init)                                   # data-dependent ActNorm initialisation happens here
    x = (1.7 * np.random.normal(size=(n, d))).astype(np.float32)
    x[0] = 3.5                                        # beyond the tail bound: identity branch of the spline
    x[1] = -3.2
    z, ldz = t.forward(x, to_numpy=True)
    xr, ldx = t.inverse(z, to_numpy=True)
    zin = (1.5 * np.random.normal(size=(n, d))).astype(np.float32)
    zin[2] = 4.0
    xo, ldo = t.inverse(zin, to_numpy=True)
    out = dict(d=d, hidden=hidden, blocks=blocks, num_bins=8, tail_bound=3, x=x, fwd_z=z, fwd_ld=ldz, inv_of_fwd_x=xr,
               inv_of_fwd_ld=ldx, zin=zin, inv_x=xo, inv_ld=ldo)
    out.update({'sd/' + k: v.detach().cpu().numpy().copy() for k, v in t.netG.state_dict().items()})
    for k in range(blocks):
        out['P/%d' % k] = t.netG.flow.flows[3 * k + 1].P.detach().cpu().numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'spline_%s.npz' % tag), **out)
    print('spline', tag, 'round trip', np.abs(xr - x).max(), 'ld sum', np.abs(ldz + ldx).max(),
          'keys', len([k for k in out if k.startswith('sd/')]))


if __name__ == '__main__':
    make_spline('d2', 2, seed=2)
    make_spline('d3', 3, seed=3)          # odd dimension: the uneven split of NSF_CL
    make_spline('d5', 5, seed=5)
    make_spline('d10', 10, seed=10)
    make_spline('d4_h8_b2', 4, hidden=8, blocks=2, seed=4)
