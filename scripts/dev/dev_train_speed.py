"""Development timing: fused epoch kernel vs the CUDA-graph autograd path, and the nearest-neighbour kernel."""
import logging
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from nnest_b200 import Trainer  # noqa: E402


def run(d, n, bs, epochs, fused):
    import os
    os.environ['NNB_TRAIN_AUTOGRAD'] = '0' if fused else '1'
    np.random.seed(0)
    torch.manual_seed(0)
    x = np.random.uniform(-1, 1, size=(n, d))
    t = Trainer(d, flow='nvp', log_dir=None, log_level=logging.WARNING, learning_rate=0.001, batch_size=bs)
    t.train(x, max_iters=2, jitter=0.01)
    torch.cuda.synchronize()
    t0 = time.time()
    t.train(x, max_iters=epochs, jitter=0.01, patience=10 ** 9)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print('d=%d n=%d bs=%d fused=%d: %.3f ms/epoch  (val loss %.5f)' % (d, n, bs, fused, 1e3 * dt / epochs,
                                                                       t.best_validation_loss))
    return t


if __name__ == '__main__':
    for d, n, bs, ep in ((2, 1000, 100, 50), (30, 65536, 100, 5), (30, 65536, 8192, 20), (50, 20000, 1000, 10)):
        run(d, n, bs, ep, True)
        run(d, n, bs, max(2, ep // 2), False)
    t = run(30, 4096, 100, 2, True)
    x = torch.rand((65536, 30), dtype=torch.float64, device='cuda')
    t.engine.mean_nn_distance(x)
    torch.cuda.synchronize()
    t0 = time.time()
    v = t.engine.mean_nn_distance(x)
    print('mean_nn_distance 65536x30: %.1f ms (%.5f)' % (1e3 * (time.time() - t0), v))
