import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) where there is no CUDA device or libnnb.so has not been built."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, 'nnest_b200', 'lib', 'libnnb.so'))
    if have_gpu and have_lib:
        return
    why = 'no CUDA device' if not have_gpu else 'libnnb.so not built (python -m nnest_b200.build)'
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
