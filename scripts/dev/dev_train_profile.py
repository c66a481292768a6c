"""Development: a few epochs of the fused fitting kernel for an ncu capture (d=30, batch 100, one CTA)."""
import logging
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from nnest_b200 import Trainer  # noqa: E402

np.random.seed(0)
torch.manual_seed(0)
x = np.random.uniform(-1, 1, size=(4096, 30))
t = Trainer(30, flow='nvp', log_dir=None, log_level=logging.WARNING, learning_rate=0.001, batch_size=100)
t.train(x, max_iters=3, jitter=0.01)
torch.cuda.synchronize()
