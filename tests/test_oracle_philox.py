"""Known-answer tests for the Philox4x32-10 stream specification (oracle/philox.py).
Vectors: Random123 kat_vectors (philox4x32 10 rounds)."""
import numpy as np

from oracle import philox


def _kat(c, k):
    return [int(x) for x in philox.philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1])]


def test_random123_known_answers():
    assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_statistics_and_independence():
    n = philox.normals(7, 3, np.arange(200000), 5)
    assert n.shape == (200000, 5) and n.dtype == np.float32
    assert abs(n.mean()) < 5e-3 and abs(n.std() - 1) < 5e-3
    assert np.abs(np.corrcoef(n.T) - np.eye(5)).max() < 1e-2
    u = philox.uniforms(7, 3, np.arange(200000))
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3
    # a chain's stream does not depend on which other chains are generated with it
    assert np.array_equal(philox.normals(7, 3, [5, 99], 5), n[[5, 99]])
    assert not np.array_equal(philox.normals(7, 4, [5], 5), n[[5]])
    assert not np.array_equal(philox.normals(8, 3, [5], 5), n[[5]])
