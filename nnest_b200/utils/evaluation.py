"""Chain diagnostics with the reference's definitions (nnest/utils/evaluation.py:6-92), vectorised over
chains and steps instead of Python double loops so that they stay usable at 10^4-10^5 chains.
Inputs are (chains, steps, dim) arrays (numpy, or torch tensors on any device)."""
import numpy as np

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _xp(x):
    return torch if (torch is not None and isinstance(x, torch.Tensor)) else np


def acceptance_rate(x):
    """Fraction of steps whose point differs from the previous one in at least one coordinate (:42-56)."""
    xp = _xp(x)
    same = (x[:, 1:] == x[:, :-1])
    same = same.all(dim=2) if xp is torch else same.all(axis=2)
    total = x.shape[0] * (x.shape[1] - 1)
    return float(total - int(same.sum())) / float(total)


def mean_jump_distance(x):
    """Mean Euclidean distance between consecutive points (:59-73)."""
    xp = _xp(x)
    dlt = x[:, 1:] - x[:, :-1]
    if xp is torch:
        return float(dlt.double().pow(2).sum(dim=2).sqrt().sum()) / (x.shape[0] * (x.shape[1] - 1))
    return float(np.sqrt((dlt.astype(np.float64) ** 2).sum(axis=2)).sum()) / (x.shape[0] * (x.shape[1] - 1))


def auto_correlation_time(x, s, mu, var):
    """Lag-s autocorrelation averaged over chains (:6-14); note the reference divides by `var` as given."""
    xp = _xp(x)
    if xp is torch:
        mu = torch.as_tensor(mu, dtype=torch.float64, device=x.device)
        var = torch.as_tensor(var, dtype=torch.float64, device=x.device)
        y = x.double() - mu
        return ((y[:, :-s] * y[:, s:]).mean(dim=1) / var).mean(dim=0).cpu().numpy()
    y = x - mu
    return ((y[:, :-s] * y[:, s:]).mean(axis=1) / var).mean(axis=0)


def effective_sample_size(x, mu, var):
    """t / (1 + 2 sum_s rho_s (1 - s/t)), lags accumulated while any rho_s > 0.05 (:17-39).  The centred series is
    formed once (float64, on the device for tensors); each lag is then one fused multiply-reduce."""
    b, t, d = x.shape
    ess = np.ones([d])
    if _xp(x) is torch:
        mu_t = torch.as_tensor(mu, dtype=torch.float64, device=x.device)
        var_t = torch.as_tensor(var, dtype=torch.float64, device=x.device)
        y = x.double() - mu_t

        def rho(s):
            return (torch.einsum('btd,btd->d', y[:, :-s], y[:, s:]) / (float(t - s) * b) / var_t).cpu().numpy()
    else:
        y = np.asarray(x, dtype=np.float64) - mu

        def rho(s):
            return np.einsum('btd,btd->d', y[:, :-s], y[:, s:]) / (float(t - s) * b) / var
    for s in range(1, t):
        p = np.asarray(rho(s))
        if np.sum(p > 0.05) == 0:
            break
        ess = ess + np.where(p > 0.05, 2.0 * p * (1.0 - float(s) / t), 0.0)
    return t / ess


def gelman_rubin_diagnostic(x, mu=None):
    """:76-92"""
    if torch is not None and isinstance(x, torch.Tensor):
        x = x.cpu().numpy()
    m, n = x.shape[0], x.shape[1]
    theta = np.mean(x, axis=1)
    sigma = np.var(x, axis=1)
    theta_m = mu if mu is not None else np.mean(theta, axis=0)
    b = float(n) / float(m - 1) * np.sum((theta - theta_m) ** 2)
    w = 1. / (float(m) * np.sum(sigma, axis=0) + 1e-5)
    v = float(n - 1) / float(n) * w + float(m + 1) / float(m * n) * b
    return np.sqrt(v / w)
