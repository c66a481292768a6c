"""Analytic likelihoods with the reference's class names and constructor signatures
(nnest/likelihoods.py), evaluated by the batched CUDA likelihood kernel (nnb_loglike) instead of a
per-row Python loop.  Each class only carries the id and parameter vector the kernel needs; calling an
instance on host arrays uploads them, runs the kernel and downloads the values -- there is no CPU
implementation in this package.
"""
import numpy as np

from . import _lib as L

_engines = {}


def _eval_engine():
    """One lazily created Engine per process/device for direct `like(x)` calls."""
    import torch
    from .engine import Engine
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if dev not in _engines:
        _engines[dev] = Engine(dev)
    return _engines[dev]


class Likelihood(object):
    num_derived = 0
    num_evaluations = 0
    like_id = None
    # True when NumPy keeps the reference's result in the input dtype (float32 rows -> float32 values)
    follows_input_dtype = False

    def __init__(self, x_dim):
        self.x_dim = x_dim

    def device_params(self):
        return []

    def __call__(self, x):
        """x: (n, x_dim) or (x_dim,) array in the likelihood's own coordinates -> log-likelihood(s)."""
        import torch
        if isinstance(x, list):
            x = np.array(x)
        x = np.asarray(x)
        single = x.ndim == 1
        xx = np.ascontiguousarray(x[None, :] if single else x)
        if xx.dtype not in (np.float32, np.float64):
            xx = xx.astype(np.float64)
        self.num_evaluations += xx.shape[0]
        eng = _eval_engine()
        eng.set_target(self.x_dim, self.like_id, self.device_params())
        out = eng.loglike(torch.from_numpy(xx).to(eng.device)).cpu().numpy()
        if self.follows_input_dtype and xx.dtype == np.float32:
            out = out.astype(np.float32)
        return out[0] if single else out

    def loglike(self, x):
        return self(np.asarray(x))

    def sample(self, prior, num_samples):
        """Rejection sampling from the likelihood under `prior` (likelihoods.py:27-36)."""
        max_loglike = self.max_loglike
        samples = np.empty((0, self.x_dim))
        while samples.shape[0] < num_samples:
            x = prior.sample(num_samples)
            ratio = np.exp(self(x) - max_loglike)
            r = np.random.uniform(low=0, high=1, size=(num_samples,))
            samples = np.vstack((x[np.where(ratio > r)], samples))
        return samples[0:num_samples]

    def uniform_sample(self, prior, num_samples, fraction):
        """Best `num_samples` of num_samples/fraction prior draws and the likelihood of the worst kept
        (likelihoods.py:38-42)."""
        x = prior.sample(int(num_samples / fraction))
        loglike = self(x)
        idx = np.argsort(-loglike)
        return x[idx[0:num_samples]], loglike[idx[num_samples - 1]]


class Rosenbrock(Likelihood):
    like_id = L.NNB_LIKE_ROSENBROCK
    follows_input_dtype = True

    @property
    def max_loglike(self):
        return self(np.ones((self.x_dim,)))

    @property
    def sample_range(self):
        return [-2] * self.x_dim, [12] * self.x_dim


class Himmelblau(Likelihood):
    like_id = L.NNB_LIKE_HIMMELBLAU
    follows_input_dtype = True
    x_dim = 2

    def __init__(self, x_dim):
        assert self.x_dim == x_dim
        super(Himmelblau, self).__init__(x_dim)

    @property
    def max_loglike(self):
        return self([3.0, 2.0])


class Gaussian(Likelihood):
    like_id = L.NNB_LIKE_GAUSSIAN

    def __init__(self, x_dim, corr, lim=5):
        self.corr = corr
        self.lim = lim
        super(Gaussian, self).__init__(x_dim)

    def device_params(self):
        return [float(self.corr)]

    @property
    def max_loglike(self):
        return self([0.0] * self.x_dim)

    @property
    def sample_range(self):
        return [-self.lim] * self.x_dim, [self.lim] * self.x_dim


class Eggbox(Likelihood):
    """(2 + prod_i cos(x_i / 2))^5.  The reference asserts x_dim == 2 (likelihoods.py:97-106); the product
    over all dimensions is the generalisation used for the x_dim=10 configuration and is identical at d=2."""
    like_id = L.NNB_LIKE_EGGBOX
    follows_input_dtype = True

    def __init__(self, x_dim=2):
        super(Eggbox, self).__init__(x_dim)

    @property
    def max_loglike(self):
        return self([0.0] * self.x_dim)


class GaussianShell(Likelihood):
    like_id = L.NNB_LIKE_GAUSSIAN_SHELL

    def __init__(self, x_dim, sigma=0.1, rshell=2, center=0):
        self.sigma = sigma
        self.rshell = rshell
        if not hasattr(center, '__len__'):
            self.center = np.array([center] * x_dim)
        else:
            self.center = np.asarray(center)
        super(GaussianShell, self).__init__(x_dim)

    def device_params(self):
        return [float(self.sigma), float(self.rshell)] + [float(c) for c in self.center]

    @property
    def max_loglike(self):
        return self(self.center - np.array([self.rshell] + [0] * (self.x_dim - 1)))


class GaussianMix(Likelihood):
    like_id = L.NNB_LIKE_GAUSSIAN_MIX

    def __init__(self, x_dim, sep=4, weights=(0.4, 0.3, 0.2, 0.1), sigma=1):
        assert len(weights) in [2, 3, 4], ('Weights must have 2, 3 or 4 components. Weights=' + str(weights))
        assert np.isclose(sum(weights), 1), ('Weights must sum to 1! Weights=' + str(weights))
        self.sep = sep
        self.weights = weights
        self.sigma = sigma
        self.sigmas = [sigma] * len(weights)
        pos = [np.asarray([0, sep]), np.asarray([0, -sep]), np.asarray([sep, 0]), np.asarray([-sep, 0])]
        self.positions = pos[:len(weights)]
        super(GaussianMix, self).__init__(x_dim)

    def device_params(self):
        return [float(self.sep), float(self.sigma), float(len(self.weights))] + [float(w) for w in self.weights]

    @property
    def max_loglike(self):
        p = np.zeros(self.x_dim)
        p[:2] = self.positions[int(np.argmax(self.weights))]
        return self(p)
