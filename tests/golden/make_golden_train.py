"""Golden fixtures for flow FITTING, recorded from the REAL reference (adammoss/nnest): its Trainer._train and
Trainer._validate (nnest/trainer.py:384-418) on its own SingleSpeedNVP + torch.optim.Adam (trainer.py:119-120).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_train.py

train_<tag>.npz: initial netG.state_dict(), x_train in visiting order (DataLoader shuffle=False), batch size, the
N(0,1) draws torch.randn_like produced for the jitter (recorded by wrapping it), learning rate / weight decay, then
  * grad/...   gradient of the first mini-batch's loss (autograd through the reference network)
  * after/...  netG.state_dict() after ONE epoch (ceil(n / batch) Adam steps)
  * after2/... after a SECOND epoch over the same order (optimizer state carried over)
  * train_loss, train_loss2, val_loss (as returned by the reference, i.e. divided by the dataset sizes)
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


def sd_arrays(netG, prefix):
    return {prefix + '/' + k: v.detach().cpu().numpy().copy() for k, v in netG.state_dict().items()}


def make_train(tag, d, n, n_valid, batch, hidden=16, layers=1, blocks=3, jitter=0.0, lr=0.001, seed=0, perturb=0.1):
    torch.manual_seed(seed)
    np.random.seed(seed)
    t = Trainer(d, hidden_dim=hidden, num_layers=layers, num_blocks=blocks, flow='nvp', batch_size=batch,
                learning_rate=lr, log_dir=None, log_level=logging.WARNING)
    with torch.no_grad():
        for p in t.netG.parameters():
            p.add_(perturb * torch.randn_like(p))
    # correlated, shifted data so that every parameter receives a non-trivial gradient
    a = np.random.normal(size=(d, d)) * 0.4 + np.eye(d)
    x = (np.random.normal(size=(n + n_valid, d)) @ a + 0.3).astype(np.float32)
    x_train, x_valid = x[:n], x[n:]
    out = dict(d=d, hidden=hidden, layers=layers, blocks=blocks, batch=batch, jitter=jitter, lr=lr,
               weight_decay=t.optimizer.defaults['weight_decay'], betas=np.array(t.optimizer.defaults['betas']),
               eps=t.optimizer.defaults['eps'], x_train=x_train, x_valid=x_valid)
    out.update(sd_arrays(t.netG, 'sd'))
    # gradient of the first mini-batch
    t.netG.zero_grad()
    loss = -t.netG.log_probs(torch.from_numpy(x_train[:batch])).mean()
    loss.backward()
    for k, p in t.netG.named_parameters():
        out['grad/' + k] = p.grad.detach().numpy().copy()
    out['first_loss'] = float(loss)
    t.netG.zero_grad()
    # two epochs through the reference's own _train, recording the jitter noise
    noise = []
    real_randn_like = torch.randn_like

    def recording(tn, *a, **k):
        r = real_randn_like(tn, *a, **k)
        noise.append(r.numpy().copy())
        return r

    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.from_numpy(x_train)), batch_size=batch,
                                         shuffle=False)
    vloader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.from_numpy(x_valid)),
                                          batch_size=x_valid.shape[0], shuffle=False, drop_last=True)
    torch.randn_like = recording
    try:
        out['train_loss'] = t._train(1, loader, jitter=jitter)
        out.update(sd_arrays(t.netG, 'after'))
        out['val_loss'] = t._validate(1, vloader)
        out['noise'] = np.concatenate(noise)
        noise[:] = []
        out['train_loss2'] = t._train(2, loader, jitter=jitter)
        out['noise2'] = np.concatenate(noise)
        out.update(sd_arrays(t.netG, 'after2'))
        out['val_loss2'] = t._validate(2, vloader)
    finally:
        torch.randn_like = real_randn_like
    np.savez_compressed(os.path.join(HERE, 'train_%s.npz' % tag), **out)
    print('train', tag, 'loss', out['first_loss'], out['train_loss'], out['train_loss2'], 'val', out['val_loss'],
          out['val_loss2'])


if __name__ == '__main__':
    make_train('d2', 2, 250, 60, 100, seed=2)                       # short tail batch (100, 100, 50)
    make_train('d5_jit', 5, 300, 64, 100, jitter=0.05, seed=5)
    make_train('d30', 30, 384, 128, 128, seed=30)
    make_train('d10_big', 10, 1000, 200, 500, seed=10)              # several CTAs per mini-batch on the device
    make_train('d7_h32_l2_b2', 7, 200, 50, 100, hidden=32, layers=2, blocks=2, seed=7)
    make_train('d50', 50, 256, 64, 128, seed=50)
