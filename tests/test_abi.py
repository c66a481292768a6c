"""CPU-side checks of the C-ABI library: it builds/loads and exports exactly what include/nnb.h declares.
No compute call is made (there is no GPU here; the library has no CPU fallback by design)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.fixture(scope='module')
def lib():
    from nnest_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'nnb.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(nnb_[a-z_0-9]+)\s*\(', txt)))


def test_header_symbols_are_exported_and_bound(lib):
    from nnest_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), 'libnnb.so does not export %s' % n
        assert n in _lib.SYMBOLS, 'ctypes binding lacks %s' % n
    assert sorted(_lib.SYMBOLS) == names
    assert lib.nnb_abi_version() == _lib.NNB_ABI_VERSION


def test_struct_layout_matches_header(lib, tmp_path):
    """sizeof / offsetof of every ABI struct, as gcc sees include/nnb.h, equal the ctypes mirror."""
    import subprocess
    from nnest_b200 import _lib
    structs = {'nnb_target': _lib.nnb_target, 'nnb_mcmc_init_args': _lib.nnb_mcmc_init_args,
               'nnb_mcmc_args': _lib.nnb_mcmc_args, 'nnb_train_args': _lib.nnb_train_args}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nnb.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ['return 0; }']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == ctypes.sizeof(cls)
        for f, _ in cls._fields_:
            assert int(got['%s.%s' % (name, f)]) == getattr(cls, f).offset, (name, f)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    rc = lib.nnb_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b'no CUDA device' in lib.nnb_last_error(None)
    from nnest_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_consume_scan_semantics(lib):
    """nested.py:429-439 : first chain from nb on whose end point differs from its start in EVERY coordinate
    and whose end loglike beats loglstar; nb advances past every inspected chain."""
    import numpy as np
    first = np.array([[0, 0], [0, 0], [0, 0], [0, 0]], dtype=np.float32)
    last = np.array([[0, 1], [1, 1], [2, 2], [3, 3]], dtype=np.float32)
    logl = np.array([9.0, 0.5, 2.0, 3.0])
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    nb = ctypes.c_int64(0)
    r = lib.nnb_consume_scan(fp(first), fp(last), logl.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 4, 2, 1.0,
                             ctypes.byref(nb))
    assert (r, nb.value) == (2, 3)        # chain 0 moved in one coordinate only, chain 1 fails the constraint
    r = lib.nnb_consume_scan(fp(first), fp(last), logl.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 4, 2, 5.0,
                             ctypes.byref(nb))
    assert (r, nb.value) == (-1, 4)


def test_chain_text_writer_matches_reference_format(lib, tmp_path):
    """nnb_write_chain_text == the reference's `' '.join('%.5E' % v for v in row)` lines (sampler.py:494-511)."""
    import numpy as np
    from nnest_b200 import _lib
    rng = np.random.RandomState(0)
    table = np.ascontiguousarray(rng.normal(size=(70001, 5)) * 10.0 ** rng.randint(-30, 30, size=(70001, 5)))
    table[0, 0], table[1, 1], table[2, 2] = 0.0, 1e-30, -9.999995e7
    table[3, 0], table[3, 1], table[3, 2], table[4, 4] = np.inf, -np.inf, np.nan, 9.9999949999999e-5
    path = str(tmp_path / 'chain.txt')
    n = lib.nnb_write_chain_text(path.encode(), b'#weight minusloglike a b c', table.ctypes.data_as(_lib._dp),
                                 table.shape[0], table.shape[1], 0)
    txt = open(path).read()
    assert n == len(txt)
    lines = txt.splitlines()
    assert lines[0] == '#weight minusloglike a b c' and len(lines) == table.shape[0] + 1
    for i in range(table.shape[0]):
        assert lines[i + 1] == ' '.join('%.5E' % v for v in table[i]), i
    # append mode, no header
    n2 = lib.nnb_write_chain_text(path.encode(), None, table.ctypes.data_as(_lib._dp), 3, 5, 1)
    assert n2 > 0 and len(open(path).read().splitlines()) == table.shape[0] + 4


def test_chain_text_fast_formatter_is_exact(lib, tmp_path):
    """The '%.5E' spelling of nnb_write_chain_text / _rows takes a fast path (six digits from a double-double product) and
    hands anything within 1e-6 of a rounding tie to std::to_chars.  Against Python's own '%.5E' on the cases that could go
    wrong: magnitudes over the whole double range, decimal near-ties d.ddddd5e+k nudged by a few ulps, decade boundaries,
    exactly representable ties, values with few mantissa bits, zeros, subnormals, the largest double."""
    import numpy as np
    from nnest_b200 import _lib
    rng = np.random.RandomState(5)
    n = 300000
    parts = [rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-307, 308, n), rng.normal(size=n) * 10.0 ** rng.randint(-3, 4, n)]
    ties = np.array([float('%d5e%d' % (m, k - 6)) for m, k in zip(rng.randint(100000, 1000000, 60000),
                                                                 rng.randint(-290, 290, 60000))])
    decades = np.array([float('1e%d' % k) for k in range(-300, 301)] + [float('9.999995e%d' % k) for k in range(-300, 300)])
    for base in (ties, decades):
        for ulps in range(-2, 3):
            x = base.copy()
            for _ in range(abs(ulps)):
                x = np.nextafter(x, np.inf if ulps > 0 else -np.inf)
            parts.append(x)
    parts.append(np.array([100000.5, 100001.5, 999999.5, 0.5, 0.25, 1.5, 2.5, 1048576.5, 3.0, 1e22, 1e23, 5e-324, 1e-290,
                           1e290, 1.7976931348623157e308, 2.2250738585072014e-308, 0.0, -0.0, 123456.5, 1234565.0]))
    parts.append(rng.randint(-10 ** 7, 10 ** 7, n // 4).astype(np.float64) / 2.0)
    parts.append(rng.randint(0, 2 ** 53, n // 4).astype(np.float64) * 2.0 ** rng.randint(-80, 80, n // 4))
    x = np.concatenate(parts)
    x = np.ascontiguousarray(np.concatenate((x, np.ones((-len(x)) % 4))).reshape(-1, 4))
    path = str(tmp_path / 'f.txt')
    assert lib.nnb_write_chain_text(path.encode(), None, x.ctypes.data_as(_lib._dp), x.shape[0], 4, 0) > 0
    with open(path) as f:
        for i, line in enumerate(f):
            assert line[:-1] == ' '.join('%.5E' % v for v in x[i]), (i, [v.hex() for v in x[i]])
    assert i + 1 == x.shape[0]


def test_save_samples_writes_the_reference_bytes(tmp_path):
    """Sampler._save_samples hands the arrays to nnb_write_chain_rows (no staging table; rows formatted in batches while the
    previous batch is written): the file is byte for byte what the reference's '%.5E' loop writes (nnest/sampler.py:
    494-511), with header, derived columns, clipped weights, non-finite values, and across batch boundaries."""
    import types
    import numpy as np
    from nnest_b200.sampler import Sampler
    rng = np.random.RandomState(0)
    for n, d, names, nder in ((1000, 5, None, 0), (37, 3, ['a', 'b', 'c'], 0), (0, 4, None, 0), (50, 2, None, 2),
                              (600001, 1, None, 0)):
        smp = rng.normal(size=(n, d)) * 10 ** rng.uniform(-8, 8, size=(n, 1))
        lgl = rng.normal(size=n) * 100
        w = rng.uniform(size=n) * 10 ** rng.uniform(-40, 0, size=n)
        if n >= 37:
            lgl[3], lgl[4], lgl[5] = np.inf, -np.inf, np.nan
            w[6], smp[7, 0] = np.nan, -np.nan
        der = rng.normal(size=(n, nder)) if nder else None
        wmax = [max(v, 1e-30) for v in w]                      # the reference's Python max(): NaN stays NaN
        cols = [np.array(wmax).reshape(n, 1), -lgl[:, None], smp] + ([der] if nder else [])
        want = ''.join(' '.join('%.5E' % v for v in row) + '\n' for row in np.concatenate(cols, axis=1))
        if names:
            want = '#weight minusloglike ' + ' '.join(names) + '\n' + want
        out = tmp_path / ('c_%d' % n)
        out.mkdir()
        me = types.SimpleNamespace(param_names=names, logs={'chains': str(out)})
        Sampler._save_samples(me, smp, lgl, weights=w, derived_samples=der)
        assert (out / 'chain.txt').read_bytes().decode() == want
        if n == 37:          # weights=None: all ones; float32 / non-contiguous inputs are converted
            Sampler._save_samples(me, smp.astype(np.float32)[:, ::-1], lgl, outfile='c2')
            want2 = ''.join(' '.join('%.5E' % v for v in [1.0, -lgl[i]] + list(smp.astype(np.float32)[i, ::-1])) + '\n'
                            for i in range(n))
            assert (out / 'c2.txt').read_bytes().decode() == '#weight minusloglike a b c\n' + want2


def test_save_samples_reproduces_the_files_the_reference_wrote(tmp_path):
    """tests/golden/chain_files.npz holds the bytes the REAL reference's Sampler._save_samples wrote (make_golden_chain.py):
    single chain with header + derived columns + weights below min_weight + infinite loglikes, single chain without
    weights, and the multi-chain layout chain_<k>.txt (nnest/sampler.py:494-527)."""
    import os
    import types
    import numpy as np
    from nnest_b200.sampler import Sampler
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'chain_files.npz'))
    for tag in ('named', 'plain', 'multi'):
        out = tmp_path / tag
        out.mkdir()
        key = lambda k: g[tag + '/' + k] if tag + '/' + k in g.files else None
        names = key('names')
        me = types.SimpleNamespace(param_names=None if names is None else [str(s) for s in names],
                                   logs={'chains': str(out)})
        Sampler._save_samples(me, key('samples'), key('loglikes'), weights=key('weights'), derived_samples=key('derived'))
        files = sorted(k.split('/', 1)[1] for k in g.files if k.startswith(tag + '/') and k.endswith('.txt'))
        assert sorted(os.listdir(str(out))) == files
        for f in files:
            assert (out / f).read_bytes() == g[tag + '/' + f].tobytes(), (tag, f)


def test_chain_text_formatter_on_adversarial_floats(lib, tmp_path):
    """hypothesis picks the doubles (subnormals, signed zeros, infinities, NaN, huge and tiny magnitudes, values one ulp
    from a power of ten): every one is spelled as Python's '%.5E' spells it."""
    import numpy as np
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import strategies as st
    from nnest_b200 import _lib
    path = str(tmp_path / 'h.txt')

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(st.lists(st.floats(allow_nan=True, allow_infinity=True, width=64), min_size=1, max_size=64))
    def check(vals):
        x = np.array(vals, dtype=np.float64).reshape(1, -1)
        assert lib.nnb_write_chain_text(path.encode(), None, x.ctypes.data_as(_lib._dp), 1, x.shape[1], 0) > 0
        assert open(path).read() == ' '.join('%.5E' % v for v in x[0]) + '\n'

    check()
