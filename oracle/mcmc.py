"""CPU oracle: the batched latent-space MCMC step of the reference, restated in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/flow.py header for who may import this).

Restates nnest/sampler.py:
  * safe_loglike / safe_prior wrappers           sampler.py:110-163
  * Sampler._mcmc_sample, hard-constraint branch sampler.py:229-370,418-463
  * Sampler._mcmc_sample, Metropolis branch      sampler.py:372-416
for prior_volume_steps=1, num_slow=0, num_derived=0 (the only combination on the hot path,
SURVEY.md appendix A).  Numeric types follow the reference: flow state float32, likelihood in
the dtype the likelihood/transform produce, loglike container float64, `scale` a Python float.

Random draws are injected through a `noise` object so that the oracle, the real reference
(monkeypatched torch.randn_like / torch.rand) and the CUDA kernel (replay buffers) can be
driven by the very same numbers:  noise.normal(step, N, d) -> float32 (N,d) and
noise.uniform(step, N) -> float32 (N,), in the reference's draw order (normal, then uniform,
once per step; sampler.py:310,334 / :377,412).
Pinned against the real reference in tests/test_oracle_golden.py.
"""
import numpy as np

from . import flow as oflow

F32 = np.float32


class ReplayNoise(object):
    def __init__(self, normals, uniforms):
        self.normals = np.asarray(normals, dtype=F32)   # (S, N, d)
        self.uniforms = np.asarray(uniforms, dtype=F32)  # (S, N)

    def normal(self, step, n, d):
        return self.normals[step - 1]

    def uniform(self, step, n):
        return self.uniforms[step - 1]


class TorchNoise(object):
    """Draws from torch's global CPU generator exactly as the reference does."""

    def normal(self, step, n, d):
        import torch
        return torch.randn(n, d).numpy()

    def uniform(self, step, n):
        import torch
        return torch.rand(n).numpy()


class Target(object):
    """loglike o transform and prior, with the reference's wrapper semantics.

    like:       object with rows()/batch() (oracle.likelihoods)
    transform:  callable (n,d)->(n,d) or None
    prior:      oracle.likelihoods.UniformPrior or None
    transform_prior: evaluate the prior on transform(x) (sampler.py:158-161)
    rowwise:    evaluate likelihood/prior with the reference's per-row Python loops (slow, used
                for the CPU baseline) instead of the vectorised restatement (same numbers).
    """

    def __init__(self, like, transform=None, prior=None, transform_prior=True, rowwise=False):
        self.like = like
        self.transform = transform if transform is not None else (lambda x: x)
        self.prior_obj = prior
        self.transform_prior = transform_prior
        self.rowwise = rowwise
        self.total_calls = 0

    def loglike(self, x):
        # sampler.py:110-133 (num_derived == 0)
        v = self.transform(x)
        logl = self.like.rows(v) if self.rowwise else self.like.batch(v)
        self.total_calls += x.shape[0]
        logl = np.array(logl, copy=True)
        if logl.ndim == 0:
            logl = logl[None]
        with np.errstate(over='ignore'):
            logl[np.logical_not(np.isfinite(logl))] = -1e100   # -inf when logl is float32
        return logl

    def prior(self, x):
        # sampler.py:143-163
        if self.prior_obj is None:
            return np.zeros(x.shape[0], dtype=np.int64)
        v = self.transform(x) if self.transform_prior else x
        return self.prior_obj.rows(v) if self.rowwise else self.prior_obj.batch(v)


def mcmc_sample(weights, target, mcmc_steps, noise, step_size=0.0, dynamic_step_size=False,
                init_samples=None, init_loglikes=None, loglstar=None, init_z=None,
                record_internals=False, flow=None):
    """Restatement of Sampler._mcmc_sample.  Returns the reference's 6-tuple
    (samples (N,S+1,d) f32, latent (N,S+1,d) f32, derived (N,S+1,0), loglikes (N,S+1) f64, scale, ncall)
    plus, when record_internals, a dict of per-step arrays (accept masks, log-ratios, ...).

    init_z: start latents for the `init_samples is None` case (the reference draws them from
            netG.prior, sampler.py:276; injected here so every path can share them).
    """
    global oflow
    _saved_flow = oflow
    if flow is not None:                              # e.g. oracle.spline: same flow_forward / flow_inverse signatures
        oflow = flow
    try:
        return _mcmc_sample(weights, target, mcmc_steps, noise, step_size, dynamic_step_size, init_samples,
                            init_loglikes, loglstar, init_z, record_internals)
    finally:
        oflow = _saved_flow


def _mcmc_sample(weights, target, mcmc_steps, noise, step_size, dynamic_step_size, init_samples, init_loglikes, loglstar,
                 init_z, record_internals):
    d = weights.d
    if step_size <= 0.0:
        step_size = 2 / d ** 0.5                       # sampler.py:248-249
    scale = step_size
    accept = reject = 0
    ncall = 0
    internals = {'mask': [], 'mask1': [], 'log_ratio': [], 'logl_prop': [], 'scale': []}

    if init_samples is not None:                      # sampler.py:262-273
        num_chains = init_samples.shape[0]
        z, _ = oflow.flow_forward(weights, np.asarray(init_samples, dtype=F32))
        x, _ = oflow.flow_inverse(weights, z)
        if init_loglikes is None:
            logl = target.loglike(x)
            ncall += num_chains
        else:
            logl = init_loglikes
        logl_prior = target.prior(x)
    else:                                             # sampler.py:275-284 (first try; caller injects z)
        z = np.asarray(init_z, dtype=F32)
        num_chains = z.shape[0]
        x, _ = oflow.flow_inverse(weights, z)
        logl = target.loglike(x)
        ncall += num_chains
        logl_prior = target.prior(x)
        if not np.all(logl > -1e30):
            raise Exception('Could not find starting value')
    logl_prior = np.asarray(logl_prior, dtype=np.float64)

    samples, latent, loglikes = [x], [z], [np.asarray(logl)]

    for it in range(1, mcmc_steps + 1):
        x, log_det_j = oflow.flow_inverse(weights, z)            # sampler.py:295
        dz = noise.normal(it, num_chains, d) * F32(scale)        # :310 / :377 (float32 tensor * python float)
        z_prop = (z + dz).astype(F32)
        x_prop, log_det_j_prop = oflow.flow_inverse(weights, z_prop)
        internals['scale'].append(scale)

        if loglstar is not None:
            log_ratio = (log_det_j_prop - log_det_j).astype(F32)  # :326
            lp_prior = target.prior(x_prop)                       # :330
            log_ratio[np.where(lp_prior < -1e30)] = -np.inf       # :331
            u = noise.uniform(it, num_chains)                     # :334
            with np.errstate(over='ignore', invalid='ignore'):
                ratio = np.minimum(np.exp(log_ratio), F32(1))     # clamp(max=1) keeps NaN
                ratio = np.where(np.isnan(log_ratio), F32(np.nan), ratio)
            mask1 = (u < ratio)
            m = mask1.astype(F32)[:, None]
            with np.errstate(invalid='ignore'):
                z_prime = (z_prop * m + z * (1 - m)).astype(F32)  # :341
                x_prime = (x_prop * m + x * (1 - m)).astype(F32)  # :342
            mask = mask1.copy()
            logl_prior_prime = target.prior(x_prime)              # :353
            logl_prime = np.full(num_chains, logl, dtype=np.float64) if np.ndim(logl) == 0 \
                else np.array(logl, dtype=np.float64, copy=True)   # :356
            lp_full = np.full(num_chains, np.nan)
            idx = np.where(mask1)[0]
            if len(idx) > 0:                                      # :359-368
                lp = target.loglike(x_prime[idx])
                ok = np.isfinite(lp) & (lp > loglstar)
                ncall += len(idx)
                logl_prime[idx[ok]] = lp[ok]
                mask[idx[~ok]] = False
                lp_full[idx] = lp
            internals['mask1'].append(mask1)
            internals['logl_prop'].append(lp_full)
        else:
            ncall += num_chains                                   # :397
            logl_prime = target.loglike(x_prop)                   # :400
            logl_prior_prime = np.asarray(target.prior(x_prop), dtype=np.float64)
            lr1 = (log_det_j_prop - log_det_j).astype(F32)        # :402
            with np.errstate(invalid='ignore'):
                lr2 = logl_prime - logl                           # :403
                lr3 = logl_prior_prime - logl_prior               # :404
                log_ratio = lr1 + lr2 + lr3                       # :405 (torch promotion == numpy here)
            u = noise.uniform(it, num_chains)                     # :412
            with np.errstate(over='ignore', invalid='ignore'):
                ratio = np.minimum(np.exp(log_ratio), 1)
                ratio = np.where(np.isnan(log_ratio), np.nan, ratio)
            mask = (u < ratio)                                    # :414
            z_prime, x_prime = z_prop, x_prop
            internals['mask1'].append(mask.copy())
            internals['logl_prop'].append(np.asarray(logl_prime, dtype=np.float64))

        num_accepted = int(mask.sum())                            # :418
        if dynamic_step_size:                                     # :422-430
            if 2 * num_accepted > num_chains:
                accept += 1
            else:
                reject += 1
            if accept > reject:
                scale *= np.exp(1. / (1 + accept))
            if accept < reject:
                scale /= np.exp(1. / (1 + reject))

        mf = mask.astype(F32)
        with np.errstate(invalid='ignore'):
            logl = logl_prime * mf + logl * (1 - mf)              # :433
            logl_prior = np.array(logl_prior, dtype=np.float64, copy=True)
            logl_prior[mask] = np.asarray(logl_prior_prime, dtype=np.float64)[mask]   # :435
            m = mf[:, None]
            z = (z_prime * m + z * (1 - m)).astype(F32)           # :437
            x = (x_prime * m + x * (1 - m)).astype(F32)           # :438
        samples.append(x)
        latent.append(z)
        loglikes.append(np.asarray(logl))
        internals['mask'].append(mask)
        internals['log_ratio'].append(np.asarray(log_ratio, dtype=np.float64))

    samples = np.transpose(np.array(samples), axes=[1, 0, 2])      # :455-458
    latent = np.transpose(np.array(latent), axes=[1, 0, 2])
    loglikes = np.transpose(np.array(loglikes, dtype=np.float64), axes=[1, 0])
    derived = np.empty((num_chains, mcmc_steps + 1, 0))
    out = (samples, latent, derived, loglikes, scale, ncall)
    if record_internals:
        return out, {k: np.array(v) for k, v in internals.items()}
    return out
