"""log-evidence of the d-dimensional Rosenbrock likelihood (nnest/likelihoods.py:50-51) under the uniform prior on
[-lim, lim]^d (examples/nested/run.py:28-30 uses lim = 5), by transfer-matrix quadrature: the likelihood is a chain
prod_i f(x_i, x_{i+1}), so   g_{d-1}(x) = 1,   g_i(x_i) = int exp(-100 (x_{i+1} - x_i^2)^2 - (1 - x_i)^2) g_{i+1} dx_{i+1}
and Z = int g_0 dx_0 / (2 lim)^d.  Trapezoid rule on n points (the narrow direction has sigma = 0.07, step 5e-4).

    python examples/nested/analytic_rosenbrock.py 2 10 30      ->  -5.8041  -43.1084  -137.4875
"""
import sys

import numpy as np
from scipy.special import logsumexp


def rosenbrock_logz(d, lim=5.0, n=20001, rows=500):
    x = np.linspace(-lim, lim, n)
    h = x[1] - x[0]
    w = np.full(n, h)
    w[0] = w[-1] = h / 2
    logw = np.log(w)
    lg = np.zeros(n)
    for _ in range(d - 1):
        new = np.empty(n)
        for s in range(0, n, rows):
            xa = x[s:s + rows]
            new[s:s + rows] = logsumexp(-100.0 * (x[None, :] - xa[:, None] ** 2) ** 2 + (lg + logw)[None, :],
                                        axis=1) - (1 - xa) ** 2
        lg = new
    return logsumexp(lg + logw) - d * np.log(2 * lim)


if __name__ == '__main__':
    for d in [int(a) for a in sys.argv[1:]] or [2]:
        print(d, '%.4f' % rosenbrock_logz(d))
