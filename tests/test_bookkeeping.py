"""CPU tests: the batched nested-sampling bookkeeping (nnest_b200/bookkeeping.py + nnb_ns_consume in the C ABI)
against the per-iteration oracle of nested.py:269-500 -- bit for bit."""
import numpy as np
import pytest

from oracle import nested as onested


@pytest.fixture(scope='module')
def lib():
    from nnest_b200 import build, _lib
    build.build()
    return _lib.load()


def _make_batches(rng, nbatch, n, d, lo, hi, p_move=0.7, ties=False):
    out = []
    for _ in range(nbatch):
        first = rng.uniform(-1, 1, size=(n, d)).astype(np.float32)
        last = first.copy()
        moved = rng.uniform(size=n) < p_move
        last[moved] += rng.uniform(0.01, 0.1, size=(moved.sum(), d)).astype(np.float32)
        half = rng.uniform(size=n) < 0.1
        last[half, 0] = first[half, 0]                     # moved in some coordinates only: not usable
        logl = rng.uniform(lo, hi, size=n)
        if ties:
            logl = np.round(logl, 1)
        out.append((first, last, logl))
    return out


def _oracle_run(u0, logl0, tr, batches, dlogz, max_iters=10 ** 9):
    it = iter(batches)
    cur = {}

    def randint(n, size):
        cur['b'] = next(it)
        return np.zeros(size, dtype=np.int64)

    def batch_fn(init_samples, init_loglikes, loglstar):
        first, last, logl = cur['b']
        return np.stack([first, last], axis=1), np.stack([np.zeros_like(logl), logl], axis=1)

    st, au, av, al, trace = onested.run_mcmc_strategy(u0, logl0, tr, batch_fn, dlogz=dlogz, max_iters=max_iters,
                                                      mcmc_num_chains=batches[0][0].shape[0], randint=randint)
    return st, au, av, al


def _bulk_run(u0, logl0, tr, batches, dlogz, chunk, max_iters=10 ** 9):
    """The driver pattern of nnest_b200/nested.py: slow path around refills, bulk in between."""
    from nnest_b200.bookkeeping import NSBook
    nlive = u0.shape[0]
    au = np.array(u0, dtype=np.float64, copy=True)
    al = np.array(logl0, dtype=np.float64, copy=True)
    av = tr(au)
    bk = NSBook(nlive)
    it_b = iter(batches)
    accept_point, get_samples, nb = True, True, 0
    first = last = logl = None
    while bk.fraction_remain > dlogz and bk.it <= max_iters:
        if get_samples or not accept_point:
            worst = int(np.argmin(al))
            loglstar = al[worst]
            if accept_point:
                bk.evidence_update(av, al, worst)
                accept_point = False
            if get_samples:
                nb = 0
                first, last, logl = next(it_b)
            # one scan by hand (nested.py:429-439)
            found = -1
            for ib in range(nb, first.shape[0]):
                nb += 1
                get_samples = nb == first.shape[0]
                if np.all(first[ib] != last[ib]) and logl[ib] > loglstar:
                    found = ib
                    break
            if found >= 0:
                au[worst] = last[found]
                av[worst] = tr(au[worst][None, :])[0]
                al[worst] = logl[found]
                accept_point = True
                bk.shrink(np.max(al))
            continue
        nb, n_done, exhausted, finished = bk.bulk(au, av, al, tr, first, last, logl, nb, chunk, dlogz, max_iters)
        if exhausted:
            accept_point = False
        get_samples = nb == first.shape[0]
        if finished:
            break
    return bk, au, av, al


@pytest.mark.parametrize('seed,nlive,n,ties,chunk', [(0, 50, 16, False, 7), (1, 200, 64, False, 1000), (2, 64, 8, True, 5),
                                                    (3, 128, 256, True, 50), (4, 30, 4, False, 3)])
def test_bulk_bookkeeping_bit_exact(lib, seed, nlive, n, ties, chunk):
    rng = np.random.RandomState(seed)
    d = 3
    tr = lambda x: 5 * x
    u0 = rng.uniform(-1, 1, size=(nlive, d))
    logl0 = rng.uniform(-50, -10, size=nlive)
    if ties:
        logl0 = np.round(logl0, 0)
    batches = _make_batches(rng, 4000, n, d, -30, 5, ties=ties)
    st, au, av, al = _oracle_run(u0, logl0, tr, batches, dlogz=0.5)
    bk, bu, bv, bl = _bulk_run(u0, logl0, tr, batches, dlogz=0.5, chunk=chunk)
    assert bk.it == st.it and bk.logz == st.logz and bk.h == st.h
    assert bk.logvol == st.logvol and bk.fraction_remain == st.fraction_remain
    sv, sl, sw = bk.dead_points()
    assert np.array_equal(sv, np.array(st.saved_v)) and np.array_equal(sl, np.array(st.saved_logl))
    assert np.array_equal(sw, np.array(st.saved_logwt))
    assert np.array_equal(bu, au) and np.array_equal(bv, av) and np.array_equal(bl, al)


def test_consume_large_batch_matches_the_single_iteration_scan(lib):
    """nnb_ns_consume at a size where its helpers run on several threads (flags of 20 000 chains) and the live set is
    radix-sorted with ties and signed zeros: the worst slots, constraints and chains it selects are those of the
    one-iteration-at-a-time loop (np.argmin + nnb_consume_scan, nested.py:272,429-439)."""
    import ctypes as C
    rng = np.random.RandomState(11)
    nlive, n, d = 3000, 20000, 8
    logl = np.round(rng.uniform(-6, 6, size=nlive), 1)
    logl[rng.randint(0, nlive, 40)] = 0.0
    logl[rng.randint(0, nlive, 40)] = -0.0
    first = rng.uniform(-1, 1, size=(n, d)).astype(np.float32)
    last = rng.uniform(-1, 1, size=(n, d)).astype(np.float32)
    stuck = rng.uniform(size=n) < 0.2
    last[stuck, 2] = first[stuck, 2]
    ll = np.round(rng.uniform(-6, 12, size=n), 1)
    ip, dp, fp = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_float)
    P = lambda a, t: a.ctypes.data_as(t)
    for max_iters in (5000, 100000):
        worst, chain, prev = (np.empty(max_iters + 1, dtype=np.int64) for _ in range(3))
        lstar, maxl = np.empty(max_iters + 1), np.empty(max_iters + 1)
        nb, exh = C.c_int64(0), C.c_int(0)
        k = lib.nnb_ns_consume(P(logl, dp), nlive, P(first, fp), P(last, fp), P(ll, dp), n, d, C.byref(nb), max_iters,
                               P(worst, ip), P(chain, ip), P(prev, ip), P(lstar, dp), P(maxl, dp), C.byref(exh))
        # the reference loop
        cur, pos, j, writer = logl.copy(), C.c_int64(0), 0, {}
        while j < max_iters:
            w = int(np.argmin(cur))
            assert worst[j] == w and lstar[j].tobytes() == cur[w].tobytes() and prev[j] == writer.get(w, -1), j
            ib = lib.nnb_consume_scan(P(first, fp), P(last, fp), P(ll, dp), n, d, float(cur[w]), C.byref(pos))
            if ib < 0:
                break
            assert chain[j] == ib
            cur[w] = ll[ib]
            writer[w] = j
            assert maxl[j] == cur.max()
            j += 1
        assert k == j and nb.value == pos.value and bool(exh.value) == (j < max_iters)


def test_bulk_respects_iteration_limit(lib):
    rng = np.random.RandomState(9)
    tr = lambda x: 5 * x
    u0 = rng.uniform(-1, 1, size=(40, 2))
    logl0 = rng.uniform(-50, -10, size=40)
    batches = _make_batches(rng, 500, 32, 2, -30, 5)
    st, au, av, al = _oracle_run(u0, logl0, tr, batches, dlogz=1e-9, max_iters=57)
    bk, bu, bv, bl = _bulk_run(u0, logl0, tr, batches, dlogz=1e-9, chunk=1000, max_iters=57)
    assert bk.it == st.it == 58 and bk.logz == st.logz and bk.h == st.h
    assert np.array_equal(bl, al) and np.array_equal(bu, au)


def test_information_recurrence_is_bit_exact(lib):
    """nnb_ns_information == the reference's Python expression (nested.py:283) evaluated iteration by iteration."""
    import ctypes as C
    rng = np.random.RandomState(0)
    n = 20000
    zn = np.sort(rng.normal(size=n) * 30) - 100
    zp = np.concatenate(([-1e300], zn[:-1]))
    lw = zn - np.abs(rng.normal(size=n))
    lstar = rng.normal(size=n) * 50
    a = np.exp(lw - zn) * lstar
    b = np.exp(zp - zn)
    h = 0.0
    for i in range(n):
        h = (a[i] + b[i] * (h + zp[i])) - zn[i]
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    got = lib.nnb_ns_information(0.0, dp(a), dp(b), dp(zp), dp(zn), n)
    assert got == h
    assert lib.nnb_ns_information(1.5, dp(a), dp(b), dp(zp), dp(zn), 0) == 1.5


def test_samples_with_is_the_concatenation_of_dead_and_live_points():
    """NSBook.samples_with (one copy of the chunked dead points) == np.concatenate((dead points, active_v)), the array the
    reference builds at the end of a run (nested.py:487-500)."""
    from nnest_b200.bookkeeping import NSBook
    rng = np.random.RandomState(3)
    bk = NSBook(10)
    av = rng.normal(size=(10, 3))
    assert np.array_equal(bk.samples_with(av), av)
    for k in (1, 5, 2):
        bk.saved_v.append(rng.normal(size=(k, 3)))
        bk.saved_logl.append(np.zeros(k))
        bk.saved_logwt.append(np.zeros(k))
    want = np.concatenate((bk.dead_points()[0].reshape(-1, 3), av))
    assert np.array_equal(bk.samples_with(av), want)


def test_consume_is_safe_to_call_from_several_threads(lib):
    """nnb_ns_consume keeps its scratch buffers between calls per THREAD: two samplers replaying their batches from two host
    threads (ctypes releases the GIL) must get what each gets alone."""
    import ctypes as C
    import threading
    ip, dp, fp = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_float)
    P = lambda a, t: a.ctypes.data_as(t)

    def problem(seed, nlive, n, d):
        rng = np.random.RandomState(seed)
        logl = rng.uniform(-50, -10, size=nlive)
        first = rng.uniform(-1, 1, size=(n, d)).astype(np.float32)
        last = rng.uniform(-1, 1, size=(n, d)).astype(np.float32)
        ll = rng.uniform(-40, 5, size=n)
        return logl, first, last, ll

    def run(prob, reps, out):
        logl, first, last, ll = prob
        n, d = first.shape
        for _ in range(reps):
            mi = n + 1
            worst, chain, prev = (np.empty(mi + 1, dtype=np.int64) for _ in range(3))
            lstar, maxl = np.empty(mi + 1), np.empty(mi + 1)
            nb, exh = C.c_int64(0), C.c_int(0)
            k = lib.nnb_ns_consume(P(logl, dp), len(logl), P(first, fp), P(last, fp), P(ll, dp), n, d, C.byref(nb), mi,
                                   P(worst, ip), P(chain, ip), P(prev, ip), P(lstar, dp), P(maxl, dp), C.byref(exh))
            out.append((k, nb.value, exh.value, worst[:k].copy(), chain[:k].copy(), lstar[:k + 1].copy(), maxl[:k].copy()))

    probs = [problem(1, 5000, 30000, 6), problem(2, 700, 9000, 30)]
    alone = [[], []]
    for i in (0, 1):
        run(probs[i], 1, alone[i])
    together = [[], []]
    th = [threading.Thread(target=run, args=(probs[i], 6, together[i])) for i in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in (0, 1):
        want = alone[i][0]
        assert len(together[i]) == 6
        for got in together[i]:
            assert got[:3] == want[:3]
            for a, b in zip(got[3:], want[3:]):
                assert np.array_equal(a, b)
