"""CPU oracle: live-point replacement and evidence bookkeeping of the reference's nested sampler.

TEST INFRASTRUCTURE ONLY (see oracle/flow.py header for who may import this).

Restates nnest/nested.py (strategy 'mcmc', plus the generic rejection consume loop):
  * worst point / weight / evidence + information update   nested.py:272-293
  * refill trigger and chain starts                        nested.py:402-415
  * sequential consumption of a batch of chains            nested.py:429-439  (rejection: 375-385)
  * volume shrink, remaining-evidence fraction             nested.py:458-464
  * final live-point contribution                          nested.py:487-500
All arithmetic float64 / integer, exactly as the reference (np.logaddexp, np.exp scalars).
Pinned against a recorded run of the real reference in tests/test_oracle_golden.py.
"""
import numpy as np


class NSState(object):
    def __init__(self, num_live_points):
        self.nlive = num_live_points
        self.h = 0.0                                                   # nested.py:242-247
        self.logz = -1e300
        self.logvol = np.log(1.0 - np.exp(-1.0 / num_live_points))
        self.fraction_remain = 1.0
        self.it = 0
        self.saved_v = []
        self.saved_logl = []
        self.saved_logwt = []
        self.accept_point = True
        self.nb = 0
        self.get_samples = True


def iteration_head(st, active_v, active_logl):
    """nested.py:272-293.  Returns (worst, loglstar)."""
    worst = int(np.argmin(active_logl))
    logwt = st.logvol + active_logl[worst]
    loglstar = active_logl[worst]
    if st.accept_point:
        logz_new = np.logaddexp(st.logz, logwt)
        st.h = (np.exp(logwt - logz_new) * active_logl[worst]
                + np.exp(st.logz - logz_new) * (st.h + st.logz) - logz_new)
        st.logz = logz_new
        st.saved_v.append(np.array(active_v[worst], copy=True))
        st.saved_logwt.append(logwt)
        st.saved_logl.append(active_logl[worst])
        st.accept_point = False
    return worst, loglstar


def consume_mcmc(st, batch_samples, batch_loglikes, worst, loglstar,
                 active_u, active_v, active_logl, transform):
    """nested.py:429-439.  batch_samples (N,S+1,d), batch_loglikes (N,S+1).  Mutates the live set."""
    n = batch_samples.shape[0]
    for ib in range(st.nb, n):
        st.nb += 1
        st.get_samples = st.nb == n
        if np.all(batch_samples[ib, 0, :] != batch_samples[ib, -1, :]) and batch_loglikes[ib, -1] > loglstar:
            active_u[worst] = batch_samples[ib, -1, :]
            active_v[worst] = transform(active_u[worst][None, :])[0]
            active_logl[worst] = batch_loglikes[ib, -1]
            st.accept_point = True
            break


def consume_rejection(st, samples, loglikes, worst, loglstar, active_u, active_v, active_logl, transform):
    """nested.py:375-385."""
    n = samples.shape[0]
    for ib in range(st.nb, n):
        st.nb += 1
        st.get_samples = st.nb == n
        if loglikes[ib] > loglstar:
            active_u[worst] = samples[st.nb - 1, :]
            active_v[worst] = transform(active_u[worst][None, :])[0]
            active_logl[worst] = loglikes[st.nb - 1]
            st.accept_point = True
            break


def iteration_tail(st, active_logl):
    """nested.py:458-464."""
    if st.accept_point:
        st.logvol -= 1.0 / st.nlive
        logz_remain = np.max(active_logl) - st.it / st.nlive
        st.fraction_remain = np.logaddexp(st.logz, logz_remain) - st.logz
        st.it += 1


def finalize(st, active_v, active_logl):
    """nested.py:487-500.  Returns (logz, h, samples, weights, loglikes, logzerr)."""
    logvol = -len(st.saved_v) / st.nlive - np.log(st.nlive)
    logz, h = st.logz, st.h
    saved_v, saved_logwt, saved_logl = list(st.saved_v), list(st.saved_logwt), list(st.saved_logl)
    for i in range(st.nlive):
        logwt = logvol + active_logl[i]
        logz_new = np.logaddexp(logz, logwt)
        h = (np.exp(logwt - logz_new) * active_logl[i] + np.exp(logz - logz_new) * (h + logz) - logz_new)
        logz = logz_new
        saved_v.append(np.array(active_v[i]))
        saved_logwt.append(logwt)
        saved_logl.append(active_logl[i])
    samples = np.array(saved_v)
    weights = np.exp(np.array(saved_logwt) - logz)
    return logz, h, samples, weights, np.array(saved_logl), np.sqrt(h / st.nlive)


def run_mcmc_strategy(active_u, active_logl, transform, batch_fn, dlogz=0.5, max_iters=1000000,
                      mcmc_num_chains=10, randint=None):
    """Outer loop nested.py:269-485 restricted to strategy=['mcmc'] without retraining hooks.

    batch_fn(init_samples, init_loglikes, loglstar) -> (samples (N,S+1,d), loglikes (N,S+1))
    randint(nlive, size) -> chain start indices (nested.py:405).
    Returns (NSState, active_u, active_v, active_logl, trace) where trace lists (worst, ib) pairs.
    """
    nlive = active_u.shape[0]
    active_u = np.array(active_u, dtype=np.float64, copy=True)
    active_logl = np.array(active_logl, dtype=np.float64, copy=True)
    active_v = transform(active_u)
    st = NSState(nlive)
    if randint is None:
        randint = lambda n, size: np.random.randint(low=0, high=n, size=size)
    trace = []
    samples = loglikes = None
    while st.fraction_remain > dlogz and st.it <= max_iters:
        worst, loglstar = iteration_head(st, active_v, active_logl)
        if st.get_samples:
            st.nb = 0
            idx = randint(nlive, mcmc_num_chains)
            samples, loglikes = batch_fn(active_u[idx, :], active_logl[idx], loglstar)
        nb0 = st.nb
        consume_mcmc(st, samples, loglikes, worst, loglstar, active_u, active_v, active_logl, transform)
        if st.accept_point:
            trace.append((worst, st.nb - 1))
        iteration_tail(st, active_logl)
    return st, active_u, active_v, active_logl, trace
