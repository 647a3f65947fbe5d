"""MT19937 jump-ahead polynomials (moocore_b200/csrc/hvb200_mtjump.c) against plain sequential generation.

The DZ2019-MC stream is the reference's MT19937 stream (c/mt19937/mt19937.c:18-55); the device generator
cuts it into segments whose start states come from t^J mod phi(t).  Host-only checks (no GPU): the
polynomial machinery must reproduce the window of the state sequence J words ahead."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32p = ctypes.POINTER(ctypes.c_uint32)


def P(a):
    return a.ctypes.data_as(u32p)


@pytest.fixture(scope="module")
def L():
    return ctypes.CDLL(os.path.join(ROOT, "moocore_b200", "libmoocore_b200.so"))


def mt_seed(seed):
    """mt19937_seed, c/mt19937/mt19937.c:6-16"""
    x = np.zeros(624, dtype=np.uint32)
    s = seed
    for i in range(624):
        x[i] = s
        s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
    return x


def sequence(state, nblocks):
    """state words followed by `nblocks` generated blocks (mt19937_gen, c/mt19937/mt19937.c:18-55)"""
    out = [state.copy()]
    x = [int(v) for v in state]
    for _ in range(nblocks):
        for i in range(624):
            y = (x[i] & 0x80000000) | (x[(i + 1) % 624] & 0x7FFFFFFF)
            x[i] = x[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        out.append(np.array(x, dtype=np.uint32))
    return np.concatenate(out)


@pytest.mark.parametrize("J", [1, 624, 1000, 19937, 24960, 624 * 64 + 7, 624 * 256])
@pytest.mark.parametrize("seed", [42, 2**32 - 1])
def test_jump_matches_sequential_generation(L, J, seed):
    poly = np.zeros(624, dtype=np.uint32)
    out = np.zeros(624, dtype=np.uint32)
    assert L.hvb200_mtjump_poly(ctypes.c_uint64(J), P(poly)) == 0
    s0 = mt_seed(seed)
    assert L.hvb200_mtjump_apply_host(P(s0), P(poly), P(out)) == 0
    seq = sequence(s0, J // 624 + 2)
    # only bit 31 of the first word of a window belongs to the state (the twist reads nothing else of it)
    assert np.array_equal(out[1:], seq[J + 1:J + 624])
    assert (int(out[0]) >> 31) == (int(seq[J]) >> 31)


def test_doubling(L):
    J = 624 * 5 + 3
    p1 = np.zeros(624, dtype=np.uint32)
    p2 = np.zeros(624, dtype=np.uint32)
    want = np.zeros(624, dtype=np.uint32)
    assert L.hvb200_mtjump_poly(ctypes.c_uint64(J), P(p1)) == 0
    for k in range(1, 5):
        assert L.hvb200_mtjump_double(P(p1), P(p2)) == 0
        assert L.hvb200_mtjump_poly(ctypes.c_uint64(J << k), P(want)) == 0
        assert np.array_equal(p2, want)
        p1, p2 = p2.copy(), p1
