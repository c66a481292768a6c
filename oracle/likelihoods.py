"""CPU oracle: analytic likelihoods and the uniform box prior of the reference.

TEST INFRASTRUCTURE ONLY (see oracle/flow.py header for who may import this).

Restates nnest/likelihoods.py and nnest/priors.py:
  * Likelihood.__call__      likelihoods.py:14-22   (per-row Python loop -> `rows()`; the same
                                                     arithmetic vectorised over rows -> `batch()`)
  * Rosenbrock.loglike       likelihoods.py:50-51
  * Himmelblau.loglike       likelihoods.py:69-70
  * Gaussian.loglike         likelihoods.py:84-86   (scipy logpdf; closed form of SURVEY appendix D in batch())
  * Eggbox.loglike           likelihoods.py:104-106 (x_dim==2 in the reference; product over all dims here,
                                                     identical for d==2)
  * GaussianShell.loglike    likelihoods.py:126-128
  * DoubleGaussianShell      likelihoods.py:143-145
  * log_gaussian_pdf/GaussianMix.loglike likelihoods.py:153-162,182-189
  * UniformPrior.__call__    priors.py:39-43 ; sample priors.py:45-47
Numeric type follows the input exactly as NumPy >= 2 does in the reference (float32 rows stay
float32 until an np.float64 scalar is mixed in); `batch()` reproduces those promotion points.
Pinned against the real reference in tests/test_oracle_golden.py.
"""
import numpy as np


class Likelihood(object):
    num_derived = 0

    def __init__(self, x_dim):
        self.x_dim = x_dim

    def __call__(self, x):
        return self.rows(x)

    def rows(self, x):
        """The reference's evaluation strategy: one Python call per row (likelihoods.py:19)."""
        x = np.asarray(x)
        if x.ndim > 1:
            return np.array([self.row(r) for r in x])
        return self.row(x)

    def batch(self, x):
        raise NotImplementedError

    def row(self, x):
        raise NotImplementedError


class Rosenbrock(Likelihood):
    like_id = 0

    def row(self, x):
        a, b = x[:-1], x[1:]
        terms = 100.0 * (b - a ** 2.0) ** 2.0 + (1 - a) ** 2.0
        return -sum(terms)  # builtin sum: left to right, in the row's own dtype

    def batch(self, x):
        x = np.asarray(x)
        a, b = x[:, :-1], x[:, 1:]
        terms = 100.0 * (b - a * a) ** 2.0 + (1 - a) ** 2.0
        acc = np.zeros(x.shape[0], dtype=x.dtype)
        for i in range(terms.shape[1]):
            acc = acc + terms[:, i]
        return -acc

    def params(self):
        return []


class Himmelblau(Likelihood):
    like_id = 1

    def __init__(self, x_dim=2):
        assert x_dim == 2
        super(Himmelblau, self).__init__(x_dim)

    def row(self, x):
        return - (x[0] ** 2 + x[1] - 11.) ** 2 - (x[0] + x[1] ** 2 - 7.) ** 2

    def batch(self, x):
        x = np.asarray(x)
        x0, x1 = x[:, 0], x[:, 1]
        return - (x0 * x0 + x1 - 11.) ** 2 - (x0 + x1 * x1 - 7.) ** 2

    def params(self):
        return []


class Gaussian(Likelihood):
    like_id = 2

    def __init__(self, x_dim, corr, lim=5):
        self.corr = corr
        self.lim = lim
        super(Gaussian, self).__init__(x_dim)

    def row(self, x):
        from scipy.stats import multivariate_normal
        d = self.x_dim
        return multivariate_normal.logpdf(x, mean=np.zeros(d), cov=np.eye(d) + self.corr * (1 - np.eye(d)))

    def batch(self, x):
        # equicorrelated closed form; always float64 like scipy's logpdf
        x = np.asarray(x, dtype=np.float64)
        d, rho = self.x_dim, float(self.corr)
        s1 = x.sum(-1)
        s2 = (x * x).sum(-1)
        a = 1.0 - rho
        bden = 1.0 - rho + d * rho
        logdet = (d - 1) * np.log(a) + np.log(bden)
        quad = (s2 - rho * s1 * s1 / bden) / a
        return -0.5 * (quad + logdet + d * np.log(2 * np.pi))

    def params(self):
        return [float(self.corr)]


class Eggbox(Likelihood):
    like_id = 3

    def row(self, x):
        chi = np.cos(x[0] / 2.)
        for i in range(1, x.shape[0]):
            chi = chi * np.cos(x[i] / 2.)
        return (2. + chi) ** 5

    def batch(self, x):
        x = np.asarray(x)
        chi = np.cos(x[:, 0] / 2.)
        for i in range(1, x.shape[1]):
            chi = chi * np.cos(x[:, i] / 2.)
        return (2. + chi) ** 5

    def params(self):
        return []


class GaussianShell(Likelihood):
    like_id = 5

    def __init__(self, x_dim, sigma=0.1, rshell=2, center=0):
        self.sigma = sigma
        self.rshell = rshell
        self.center = np.array([center] * x_dim) if not hasattr(center, '__len__') else np.asarray(center)
        super(GaussianShell, self).__init__(x_dim)

    def row(self, x):
        rad = np.sqrt(np.sum((self.center - x) ** 2))
        return - ((rad - self.rshell) ** 2) / (2 * self.sigma ** 2)

    def batch(self, x):
        x = np.asarray(x)
        rad = np.sqrt(np.sum((self.center - x) ** 2, axis=-1))
        return - ((rad - self.rshell) ** 2) / (2 * self.sigma ** 2)

    def params(self):
        return [float(self.sigma), float(self.rshell)] + [float(c) for c in self.center]


class GaussianMix(Likelihood):
    like_id = 4

    def __init__(self, x_dim, sep=4, weights=(0.4, 0.3, 0.2, 0.1), sigma=1):
        assert len(weights) in [2, 3, 4]
        self.sep = sep
        self.weights = weights
        self.sigma = sigma
        pos = [(0, sep), (0, -sep), (sep, 0), (-sep, 0)]
        self.positions = [np.asarray(p) for p in pos[:len(weights)]]
        super(GaussianMix, self).__init__(x_dim)

    def _log_gauss(self, theta):
        # likelihoods.py:153-162 with mu=0, ndim=len(theta)
        sigma = self.sigma
        logl = -(np.sum(theta ** 2) / (2 * sigma ** 2))
        logl = logl - np.log(2 * np.pi * (sigma ** 2)) * len(theta) / 2.0
        return logl

    def row(self, theta):
        import scipy.special
        logls = []
        for k, pos in enumerate(self.positions):
            th = np.array(theta, copy=True)
            th[:2] -= pos
            logls.append(self._log_gauss(th) + np.log(self.weights[k]))
        return scipy.special.logsumexp(logls)

    def batch(self, x):
        x = np.asarray(x)
        sigma = self.sigma
        cols = []
        for k, pos in enumerate(self.positions):
            th = np.array(x, copy=True)
            th[:, :2] -= pos
            q = -(np.sum(th ** 2, axis=-1) / (2 * sigma ** 2))        # stays in x.dtype
            q = q - np.log(2 * np.pi * (sigma ** 2)) * x.shape[1] / 2.0  # np.float64 scalar -> float64
            cols.append(q + np.log(self.weights[k]))
        a = np.stack(cols, axis=-1).astype(np.float64)
        m = a.max(-1)
        return m + np.log(np.exp(a - m[:, None]).sum(-1))

    def params(self):
        return [float(self.sep), float(self.sigma), float(len(self.weights))] + [float(w) for w in self.weights]


class UniformPrior(object):
    """priors.py:24-47"""

    def __init__(self, x_dim, minimum, maximum):
        self.x_dim = x_dim
        self.minimum = np.array([minimum] * x_dim) if not hasattr(minimum, '__len__') else np.array(minimum)
        self.maximum = np.array([maximum] * x_dim) if not hasattr(maximum, '__len__') else np.array(maximum)

    def __call__(self, x):
        # one row (priors.py:39-43)
        if np.any(x < self.minimum) or np.any(x > self.maximum):
            return -np.inf
        return 0

    def rows(self, x):
        return np.array([self(r) for r in x])

    def batch(self, x):
        x = np.asarray(x)
        bad = np.any(x < self.minimum, axis=-1) | np.any(x > self.maximum, axis=-1)
        return np.where(bad, -np.inf, 0.0)

    def sample(self, num_samples, rng=np.random):
        return self.minimum + (self.maximum - self.minimum) * rng.uniform(size=(num_samples, self.x_dim))
