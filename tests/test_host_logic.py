"""CPU tests of the host-side logic that needs no GPU: transform probing, state_dict flattening, the run-directory
layout, error behaviour without a CUDA device."""
import logging
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import flow as oflow
from oracle import train as otrain
from helpers import load, state_dict_of, weights_of


def test_probe_affine_transform_recovers_the_reference_maps():
    """the maps of examples/nested/run.py:25-44 and nnest/mcmc.py:111"""
    from nnest_b200.sampler import probe_affine_transform
    for fn, sc, sh in [(lambda x: 5 * x, 5.0, 0.0), (lambda x: x * 5 * np.pi, 5 * np.pi, 0.0),
                       (lambda x: 10 * x, 10.0, 0.0)]:
        scale, shift, promotes = probe_affine_transform(fn, 4)
        assert np.allclose(scale, sc, rtol=1e-15) and np.allclose(shift, sh) and not promotes
    std, mean = np.array([0.5, 2.0, 3.0]), np.array([1.0, -1.0, 0.25])
    scale, shift, promotes = probe_affine_transform(lambda x: x * std + mean, 3)
    assert np.allclose(scale, std) and np.allclose(shift, mean)
    assert promotes                      # float64 arrays promote float32 input: the likelihood then runs in float64
    assert probe_affine_transform(None, 3) == (None, None, False)


@pytest.mark.parametrize('fn', [lambda x: x ** 2, lambda x: np.sin(x), lambda x: x @ np.array([[1.0, 0.5], [0.0, 1.0]]),
                                lambda x: x[:, :1]])
def test_probe_rejects_non_affine_or_mixing_transforms(fn):
    from nnest_b200.sampler import probe_affine_transform
    with pytest.raises(NotImplementedError):
        probe_affine_transform(fn, 2)


@pytest.mark.parametrize('name', ['flow_d5.npz', 'flow_d7_h32_l2_b5.npz', 'flow_d6_translate.npz', 'flow_d6_constant.npz'])
def test_flatten_state_dict_matches_the_oracle_layout(name):
    """engine.flatten_state_dict (the nnb_set_flow input) == the oracle's own flattening of the same golden weights."""
    from nnest_b200.engine import flatten_state_dict
    g = load(name)
    scale = str(g['scale']) if g['scale'].dtype.kind in 'US' else ''
    flat, d, hidden, nl, nb, flags = flatten_state_dict(state_dict_of(g), scale)
    w = weights_of(g)
    assert (d, hidden, nl, nb) == (w.d, w.hidden, w.num_layers, w.num_blocks)
    assert np.array_equal(flat, w.flat())
    if scale == '':
        assert np.array_equal(flat, otrain.flatten_state_dict(state_dict_of(g), nb).astype(np.float32))
        assert flat.size == 2 * nb * otrain.net_floats(d, hidden, nl)


def test_run_directory_layout(tmp_path):
    """<log_dir>/runN/{info,results,chains,checkpoint,plots} (nnest/utils/logger.py:38-75)"""
    from nnest_b200.utils.logger import get_or_create_run_dir
    a = get_or_create_run_dir(str(tmp_path / 'logs'))
    b = get_or_create_run_dir(str(tmp_path / 'logs'))
    assert a['created'] and b['created']
    assert a['run_dir'].endswith('run1') and b['run_dir'].endswith('run2')
    for sub in ('info', 'results', 'chains', 'checkpoint', 'plots'):
        assert os.path.isdir(a[sub])
    c = get_or_create_run_dir(a['run_dir'])              # an existing run directory is resumed, not renumbered
    assert not c['created'] and c['run_dir'] == a['run_dir']
    d = get_or_create_run_dir(str(tmp_path / 'flat'), append_run_num=False)
    assert d['created'] and d['run_dir'].endswith('flat')


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour WITHOUT a CUDA device')
def test_samplers_and_trainer_refuse_to_run_without_a_gpu(tmp_path):
    from nnest_b200 import NestedSampler, Trainer
    from nnest_b200.likelihoods import Rosenbrock
    with pytest.raises(Exception):
        Trainer(2, flow='nvp', log_dir=None, log_level=logging.ERROR)
    with pytest.raises(Exception):
        NestedSampler(2, Rosenbrock(2), flow='nvp', log_dir=str(tmp_path), log_level=logging.ERROR)
    with pytest.raises(NotImplementedError):
        Trainer(2, flow='choleksy', log_dir=None)          # only 'nvp' and 'spline' are implemented on the device
    with pytest.raises(NotImplementedError):
        Trainer(2, flow='spline', num_slow=1, log_dir=None)
