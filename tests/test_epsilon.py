"""Epsilon indicator (c/epsilon.h:52-221): oracle port vs the compiled reference and the reference's known answers
(CPU), GPU entry points vs the oracle bit for bit (GPU)."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = ctypes.POINTER(ctypes.c_double)
u8p = ctypes.POINTER(ctypes.c_uint8)


def P(a):
    return a.ctypes.data_as(dp)


def orc_eps(oracle, mult, a, b, maxi):
    a = np.ascontiguousarray(a, dtype=float)
    b = np.ascontiguousarray(np.atleast_2d(b), dtype=float)
    mx = np.ascontiguousarray(np.broadcast_to(np.asarray(maxi, bool), (a.shape[1],))).view(np.uint8)
    oracle.orc_epsilon.restype = ctypes.c_double
    return oracle.orc_epsilon(int(mult), P(a), ctypes.c_size_t(a.shape[0]), a.shape[1], P(b), ctypes.c_size_t(b.shape[0]),
                              mx.ctypes.data_as(u8p))


def ref_eps(L, mult, a, b, maxi):
    a = np.ascontiguousarray(a, dtype=float)
    b = np.ascontiguousarray(np.atleast_2d(b), dtype=float)
    mx = np.ascontiguousarray(np.broadcast_to(np.asarray(maxi, bool), (a.shape[1],))).view(np.uint8)
    fn = L.ref_epsilon_mult if mult else L.ref_epsilon_additive
    fn.restype = ctypes.c_double
    return fn(P(a), ctypes.c_size_t(a.shape[0]), a.shape[1], P(b), ctypes.c_size_t(b.shape[0]), mx.ctypes.data_as(u8p))


# known answers of the reference's own tests: python/tests/test_moocore.py:386-397, r/tests/testthat/test-doctest-epsilon_additive.R:7-13,
# python/src/moocore/_moocore.py:495-502
REF10 = np.array([10, 1, 6, 1, 2, 2, 1, 6, 1, 10], dtype=float).reshape((-1, 2))
A3 = np.array([4, 2, 3, 3, 2, 4], dtype=float).reshape((-1, 2))
RA1 = np.array([9, 2, 8, 4, 7, 5, 5, 6, 4, 7], dtype=float).reshape((-1, 2))
RA2 = np.array([8, 4, 7, 5, 5, 6, 4, 7], dtype=float).reshape((-1, 2))
RA3 = np.array([10, 4, 9, 5, 8, 6, 7, 7, 6, 8], dtype=float).reshape((-1, 2))
DOC_DAT = np.array([[3.5, 5.5], [3.6, 4.1], [4.1, 3.2], [5.5, 1.5]])
DOC_REF = np.array([[1, 6], [2, 5], [3, 4], [4, 3], [5, 2], [6, 1]], dtype=float)
KATS = [
    (False, A3, REF10, False, 1.0), (True, A3, REF10, False, 2.0), (True, A3, REF10, True, 2.5), (False, A3, REF10, True, 6.0),
    (False, -A3, -REF10, True, 1.0), (False, -A3, -REF10, False, 6.0), (True, A3, REF10, [True, False], 2.5),
    (True, A3, REF10, [False, True], 2.5), (True, RA1, RA3, False, 0.9), (True, RA1, RA2, False, 1.0), (True, RA2, RA1, False, 2.0),
    (False, DOC_DAT, DOC_REF, False, 2.5), (True, DOC_DAT, DOC_REF, False, 3.5),
]


def random_cases():
    rng = np.random.default_rng(12)
    for i in range(40):
        d = int(rng.integers(2, 13))
        na, nb = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        a = rng.random((na, d)) + 0.05
        b = rng.random((nb, d)) + 0.05
        if i % 5 == 0:
            b[: min(na, nb)] = a[: min(na, nb)]              # ties
        maxi = rng.random(d) < 0.4 if i % 3 else bool(i % 2)
        yield a, b, maxi


@pytest.mark.parametrize("case", range(len(KATS)))
def test_oracle_known_answers(oracle, case):
    mult, a, b, maxi, want = KATS[case]
    assert orc_eps(oracle, mult, a, b, maxi) == pytest.approx(want, rel=1e-15)


def test_oracle_port_equals_compiled_reference(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libepsref.so")
    if not os.path.exists(path):
        pytest.skip("compiled reference not present (needs /root/reference at build time)")
    L = ctypes.CDLL(path)
    for mult, a, b, maxi, want in KATS:
        assert ref_eps(L, mult, a, b, maxi) == pytest.approx(want, rel=1e-15)
    for a, b, maxi in random_cases():
        for mult in (False, True):
            assert orc_eps(oracle, mult, a, b, maxi) == ref_eps(L, mult, a, b, maxi)
        assert orc_eps(oracle, False, -a, b - 2.0, maxi) == ref_eps(L, False, -a, b - 2.0, maxi)       # negative values


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(KATS)))
def test_gpu_known_answers(mb, case):
    mult, a, b, maxi, want = KATS[case]
    got = (mb.epsilon_mult if mult else mb.epsilon_additive)(a, b, maximise=maxi)
    assert got == pytest.approx(want, rel=1e-15)


@pytest.mark.gpu
def test_gpu_bit_exact_vs_oracle(mb, oracle):
    for a, b, maxi in random_cases():
        assert mb.epsilon_additive(a, b, maximise=maxi) == orc_eps(oracle, False, a, b, maxi)
        assert mb.epsilon_mult(a, b, maximise=maxi) == orc_eps(oracle, True, a, b, maxi)
        assert mb.epsilon_additive(-a, b - 2.0, maximise=maxi) == orc_eps(oracle, False, -a, b - 2.0, maxi)
    rng = np.random.default_rng(1)
    a, b = rng.random((5000, 10)) + 0.1, rng.random((3001, 10)) + 0.1      # several tiles and blocks
    assert mb.epsilon_additive(a, b) == orc_eps(oracle, False, a, b, False)
    assert mb.epsilon_mult(a, b, maximise=[True] * 3 + [False] * 7) == orc_eps(oracle, True, a, b, [True] * 3 + [False] * 7)
    a, b = rng.random((70, 40)) + 0.1, rng.random((90, 40)) + 0.1
    assert mb.epsilon_mult(a, b) == orc_eps(oracle, True, a, b, False)


@pytest.mark.gpu
def test_gpu_argument_errors(mb):
    with pytest.raises(ValueError):
        mb.epsilon_mult(np.array([[1.0, -1.0]]), np.array([[1.0, 1.0]]))
    with pytest.raises(ValueError):
        mb.epsilon_mult(np.array([[1.0, 1.0]]), np.array([[0.0, 1.0]]))
    with pytest.raises(ValueError):
        mb.epsilon_additive(np.zeros((2, 3)), np.zeros((2, 2)))
    assert mb.epsilon_additive(np.array([[1.0, 2.0]]), [0.5, 0.5]) == 1.5             # a single reference point as a vector
