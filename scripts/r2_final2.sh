cd $GRAFT_REPO_ROOT
bash scripts/final_check.sh
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
from nnest_b200.engine import Engine
eng = Engine(0)
x = torch.from_numpy(np.random.RandomState(0).uniform(-1, 1, size=(65536, 30))).cuda()
for _ in range(2): v = eng.mean_nn_distance(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): v = eng.mean_nn_distance(x)
torch.cuda.synchronize(); print('mean_nn_distance 65536 x 30: %.2f ms (value %.6f)' % (1e3 * (time.perf_counter() - t0) / 5, v))
PY
bash scripts/r2_ns_full.sh
