"""Philox4x32-10 oracle against the published known-answer vectors (Random123 kat_vectors,
'philox4x32 10' rows) - the one part of the oracle with an external golden source."""
import numpy as np

from oracle.philox import philox4x32, uniform01, draws

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_published_vectors():
    for ctr, key, want in KAT:
        got = philox4x32(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == want


def test_batched_equals_scalar_and_uniform_range():
    ids = np.arange(1000, dtype=np.uint32)
    d = draws(12345678901234, ids, 17, 2)
    assert d.shape == (1000, 4) and d.dtype == np.uint32
    key = np.array([12345678901234 & 0xFFFFFFFF, 12345678901234 >> 32], dtype=np.uint32)
    for i in (0, 1, 999):
        assert np.array_equal(d[i], philox4x32(np.array([i, 17, 2, 0], dtype=np.uint32), key))
    u = uniform01(d)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert uniform01(np.uint32(0xFFFFFFFF)) == np.float32(1.0 - 2.0 ** -24)
    assert abs(float(u.mean()) - 0.5) < 0.02


def test_streams_and_steps_are_independent():
    ids = np.arange(64, dtype=np.uint32)
    a, b, c = draws(1, ids, 0, 0), draws(1, ids, 1, 0), draws(1, ids, 0, 1)
    assert not np.array_equal(a, b) and not np.array_equal(a, c) and not np.array_equal(b, c)
    # sharding invariance: the draw depends on the GLOBAL env id only
    assert np.array_equal(draws(1, ids[32:], 5, 0), draws(1, ids, 5, 0)[32:])
