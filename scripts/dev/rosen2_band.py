"""logZ of the reference's integration test setting (tests/test_nested.py:10-19, flow='nvp') over a few seeds: how far
inside the reference's +-0.2 band does the package sit?"""
import logging, sys, tempfile
import numpy as np
import torch
sys.path.insert(0, '.')
from nnest_b200 import NestedSampler
from nnest_b200.likelihoods import Rosenbrock
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    np.random.seed(seed)
    torch.manual_seed(seed)
    s = NestedSampler(2, Rosenbrock(2), transform=lambda x: 5 * x, num_live_points=1000, hidden_dim=16, num_layers=1,
                      num_blocks=3, num_slow=0, flow='nvp', log_dir=tempfile.mkdtemp(), log_level=logging.WARNING, seed=seed)
    s.run(mcmc_num_chains=1000, mcmc_dynamic_step_size=False, train_iters=200)
    print('seed %d logz %.4f +- %.4f  |logz + 5.80| = %.4f' % (seed, s.logz, s.logzerr, abs(s.logz + 5.80)), flush=True)
