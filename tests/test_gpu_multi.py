"""Multi-GPU test (needs >= 2 CUDA devices; skipped on the single-GPU tier): chain sharding over NCCL."""
import json
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharding_and_nested_run(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = tmp_path / 'result.json'
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'multi_gpu_worker.py')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), worker, str(out)]
    subprocess.run(cmd, check=True, timeout=900)
    res = json.load(open(out))
    assert res['world'] == 2 and res['shard_ok'] and res['ranks_identical']
    assert abs(res['logz'] + 5.804) < 0.4 + 3 * res['logzerr']
    assert res['default_strategy_ranks_identical']
    assert abs(res['default_strategy_logz'] + 5.804) < 0.4 + 3 * res['default_strategy_logzerr']
    (tmp_path / 'ok').write_text(json.dumps(res))
    print('MULTI-GPU RESULT', json.dumps(res))
