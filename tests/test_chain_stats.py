"""Chain diagnostics against the golden recorded from the reference's nnest/utils/evaluation.py
(tests/golden/make_golden_stats.py): the vectorised NumPy restatement (CPU) and the CUDA kernels (GPU)."""
import numpy as np
import pytest

from helpers import load

CASES = ['a', 'b', 'c']


@pytest.mark.parametrize('tag', CASES)
def test_numpy_restatement_matches_reference(tag):
    from nnest_b200.utils.evaluation import acceptance_rate, effective_sample_size, mean_jump_distance
    g = load('chain_stats.npz')
    v = g[tag + '_x'] * g[tag + '_scale'] + g[tag + '_shift']
    assert acceptance_rate(v) == float(g[tag + '_acceptance'])
    assert np.allclose(effective_sample_size(v, g[tag + '_mean'], g[tag + '_std']), g[tag + '_ess'], rtol=1e-10)
    assert abs(mean_jump_distance(v) - float(g[tag + '_jump'])) <= 1e-12 * float(g[tag + '_jump'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', CASES)
def test_device_kernels_match_reference(tag):
    import torch
    from nnest_b200.engine import Engine
    eng = Engine(0)
    g = load('chain_stats.npz')
    x = g[tag + '_x']                                                        # (b, t, d)
    trace = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 2, 0))).cuda()   # [T][d][N]
    acc, ess, jump = eng.chain_stats(trace, t_scale=g[tag + '_scale'], t_shift=g[tag + '_shift'])
    assert acc == float(g[tag + '_acceptance'])
    assert np.allclose(ess, g[tag + '_ess'], rtol=1e-9)
    assert abs(jump - float(g[tag + '_jump'])) <= 1e-9 * float(g[tag + '_jump'])
    # explicit mean / std (the nested-sampling log line passes those of the live points) and a prefix of the trace
    t = x.shape[1] // 2
    v = x[:, :t + 1] * g[tag + '_scale'] + g[tag + '_shift']
    from nnest_b200.utils.evaluation import acceptance_rate, effective_sample_size, mean_jump_distance
    mean, std = g[tag + '_mean'] + 0.1, g[tag + '_std'] * 1.3
    acc, ess, jump = eng.chain_stats(trace, steps=t, t_scale=g[tag + '_scale'], t_shift=g[tag + '_shift'], mean=mean,
                                     std=std)
    assert acc == acceptance_rate(v)
    assert np.allclose(ess, effective_sample_size(v, mean, std), rtol=1e-9)
    assert abs(jump - mean_jump_distance(v)) <= 1e-9 * jump
