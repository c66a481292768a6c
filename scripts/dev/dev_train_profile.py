"""Development: a few epochs of the fused fitting kernel for an ncu capture.
usage: python scripts/dev/dev_train_profile.py [n_samples=65536] [batch_size=8192]   (the C4 retrain shape: d = 30)"""
import logging
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from nnest_b200 import Trainer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
np.random.seed(0)
torch.manual_seed(0)
x = np.random.uniform(-1, 1, size=(n, 30))
t = Trainer(30, flow='nvp', log_dir=None, log_level=logging.WARNING, learning_rate=0.001, batch_size=bs)
t.train(x, max_iters=5, jitter=0.01)
torch.cuda.synchronize()
print('epochs done', t.total_iters, 'samples', n, 'batch', bs)
