"""Priors with the reference's interface (nnest/priors.py): a Prior is called on ONE point and returns its
log-density; UniformPrior returns 0 inside the box and -inf outside, and can draw uniform samples.

Host-side helpers (seeding the live points, reference nnest/nested.py:207).  On the hot path the box test
runs inside the CUDA kernels (nnb_set_target prior_kind / prior_lo / prior_hi); Sampler recognises
UniformPrior and ships its bounds to the device."""
import numpy as np


class Prior(object):

    def __init__(self, x_dim):
        self.x_dim = x_dim

    def __call__(self, x):
        x = np.asarray(x)
        if x.ndim > 1:
            return np.array([self.loglike(r) for r in x])
        return self.loglike(x)

    def loglike(self, x):
        raise NotImplementedError

    def sample(self, num_samples):
        raise NotImplementedError


class UniformPrior(Prior):

    def __init__(self, x_dim, minimum, maximum):
        super(UniformPrior, self).__init__(x_dim)
        self.minimum = self._bounds(minimum, x_dim)
        self.maximum = self._bounds(maximum, x_dim)

    @staticmethod
    def _bounds(v, x_dim):
        if hasattr(v, '__len__'):
            assert len(v) == x_dim
            return np.array(v)
        return np.array([v] * x_dim)

    def __call__(self, x):
        outside = np.any(x < self.minimum) or np.any(x > self.maximum)
        return -np.inf if outside else 0

    def sample(self, num_samples):
        width = self.maximum - self.minimum
        return self.minimum + width * np.random.uniform(size=(num_samples, self.x_dim))
