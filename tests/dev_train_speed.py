import sys, os, time, logging
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from nnest_b200 import Trainer
np.random.seed(0); torch.manual_seed(0)
for d, n, bs in ((2, 1000, 100), (30, 16384, 1024)):
    t = Trainer(d, flow='nvp', log_dir=None, learning_rate=0.001, batch_size=bs, log_level=logging.WARNING)
    x = np.random.normal(size=(n, d)) * 0.1 + 0.2
    before = -t.log_probs(x.astype(np.float32)).mean().item()
    torch.cuda.synchronize(); t0 = time.time()
    t.train(x, max_iters=50, jitter=0.01)
    torch.cuda.synchronize(); dt = time.time() - t0
    after = -t.log_probs(x.astype(np.float32)).mean().item()
    print('d=%d n=%d bs=%d: 50 epochs %.2f s (%.2f ms/iteration)  loss %.3f -> %.3f' % (d, n, bs, dt, 1e3 * dt / (50 * ((n * 9 // 10) // bs + 1)), before, after))
