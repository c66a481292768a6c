"""CPU: the neural-spline-flow oracle (oracle/spline.py) against goldens recorded from the real reference
(tests/golden/make_golden_spline.py).  Groundwork for SURVEY section 8(f) #3; tolerance 1e-5 relative like the NVP flow,
2e-5 absolute for log-determinants (sums of ~3 d logarithms in float32)."""
import numpy as np
import pytest

from oracle import spline as ospline
from helpers import load, rel_err

CASES = ['d2', 'd3', 'd5', 'd10', 'd4_h8_b2']


@pytest.mark.parametrize('name', CASES)
def test_spline_flow_matches_reference(name):
    g = load('spline_%s.npz' % name)
    w = ospline.SplineWeights.from_golden(g)
    z, ld = ospline.flow_forward(w, g['x'])
    assert rel_err(z, g['fwd_z']) < 1e-5
    assert np.abs(ld - g['fwd_ld']).max() < 2e-5 * max(1.0, np.abs(g['fwd_ld']).max())
    x, ldx = ospline.flow_inverse(w, g['zin'])
    assert rel_err(x, g['inv_x']) < 1e-5
    assert np.abs(ldx - g['inv_ld']).max() < 2e-5 * max(1.0, np.abs(g['inv_ld']).max())
    # the properties tests/test_flows.py checks for every flow: round trip and log-det antisymmetry
    xr, ldr = ospline.flow_inverse(w, z)
    assert np.abs(xr - g['x']).max() < 2e-5
    assert np.abs(ld + ldr).max() < 5e-5


def test_identity_outside_the_tail_bound():
    g = load('spline_d2.npz')
    w = ospline.SplineWeights.from_golden(g)
    blk = w.blocks[0]
    x = np.array([[3.5, -3.2], [0.3, 4.0], [0.1, -0.4]], dtype=np.float32)
    y, ld = ospline._coupling(w, blk, x, False)
    assert y[0, 0] == x[0, 0] and y[0, 1] == x[0, 1] and ld[0] == 0.0      # both halves outside: untouched
    assert y[1, 1] == x[1, 1] and y[1, 0] != x[1, 0]
    # a batch with NO coordinate inside the interval raises like the reference (RQS on an empty selection,
    # networks.py:464-465); Sampler._mcmc_sample turns that into a skipped proposal (sampler.py:320-324)
    with pytest.raises(ValueError):
        ospline._coupling(w, blk, x[:1], False)


@pytest.mark.parametrize('name', CASES)
def test_kernel_header_arithmetic_matches_reference(name):
    """nnest_b200/csrc/nnb_spline.cuh (the __host__ __device__ per-sample code a CUDA kernel will wrap), compiled for the
    CPU by oracle/build_spline_host.py, against the same goldens."""
    import ctypes as C
    from oracle import build_spline_host
    lib = C.CDLL(build_spline_host.build())
    g = load('spline_%s.npz' % name)
    w = ospline.SplineWeights.from_golden(g)
    d, hidden, blocks = int(g['d']), int(g['hidden']), int(g['blocks'])
    packed = np.ascontiguousarray(ospline.pack_for_kernel(w, hidden))
    assert packed.size == blocks * lib.spline_host_block_floats(d, hidden, w.K)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))

    def run(inp, inverse):
        inp = np.ascontiguousarray(inp, dtype=np.float32)
        out, ld = np.empty_like(inp), np.empty(inp.shape[0], dtype=np.float32)
        rc = lib.spline_host_flow(fp(packed), d, hidden, blocks, w.K, C.c_float(w.B), int(inverse), fp(inp), fp(out), fp(ld),
                                  C.c_int64(inp.shape[0]))
        assert rc == 0
        return out, ld

    z, ld = run(g['x'], False)
    assert rel_err(z, g['fwd_z']) < 1e-5
    assert np.abs(ld - g['fwd_ld']).max() < 2e-5 * max(1.0, np.abs(g['fwd_ld']).max())
    x, ldx = run(g['zin'], True)
    assert rel_err(x, g['inv_x']) < 1e-5
    assert np.abs(ldx - g['inv_ld']).max() < 2e-5 * max(1.0, np.abs(g['inv_ld']).max())
    xr, ldr = run(z, True)
    assert np.abs(xr - g['x']).max() < 2e-5 and np.abs(ld + ldr).max() < 5e-5
