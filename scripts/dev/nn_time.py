"""Development probe: time of nnb_mean_nn_distance at the config-4 retrain size (65 536 x 30, float64 rows)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nnest_b200.engine import Engine
eng = Engine(0)
shapes = ((65536, 30), (16384, 10), (65536, 50)) if not os.environ.get('NNB_NN_ONE') else ((65536, 30),)
for n, d in shapes:
    x = torch.from_numpy(np.random.RandomState(0).uniform(-1, 1, size=(n, d))).cuda()
    for _ in range(2):
        v = eng.mean_nn_distance(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        v = eng.mean_nn_distance(x)
    torch.cuda.synchronize()
    print('%s QT=%s ST=%s mean_nn_distance %d x %d: %.2f ms (value %.15g)' % ('f32' if os.environ.get('NNB_NN_NO_TC') else 'tc ', os.environ.get('NNB_NN_QT'), os.environ.get('NNB_NN_STAGES'), n, d, 1e3 * (time.perf_counter() - t0) / 5, v))
