"""nnest_b200 -- B200-native (sm_100a) implementation of the batched latent-space MCMC hot path of
adammoss/nnest, behind the reference's own Python API (NestedSampler / MCMCSampler / Trainer).

Importing the package is cheap and works without a GPU; constructing a sampler/trainer/engine needs libnnb.so
(python -m nnest_b200.build) and a CUDA device -- there is no CPU fallback.
"""
__version__ = '0.1.0'

_LAZY = {
    'NestedSampler': ('nested', 'NestedSampler'),
    'MCMCSampler': ('mcmc', 'MCMCSampler'),
    'Sampler': ('sampler', 'Sampler'),
    'Trainer': ('trainer', 'Trainer'),
    'Engine': ('engine', 'Engine'),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module('.' + _LAZY[name][0], __name__)
        return getattr(mod, _LAZY[name][1])
    raise AttributeError(name)
