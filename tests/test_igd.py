"""IGD, IGD+ and the averaged Hausdorff distance (c/igd.h:60-289): oracle port vs the compiled reference and the
reference's known answers (CPU); GPU entry points vs the oracle bit for bit (GPU)."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = ctypes.POINTER(ctypes.c_double)
u8p = ctypes.POINTER(ctypes.c_uint8)


def P(a):
    return a.ctypes.data_as(dp)


def _prep(a, b, maxi):
    a = np.ascontiguousarray(a, dtype=float)
    b = np.ascontiguousarray(np.atleast_2d(b), dtype=float)
    mx = np.ascontiguousarray(np.broadcast_to(np.asarray(maxi, bool), (a.shape[1],))).view(np.uint8)
    return a, b, mx


def orc(oracle, what, a, b, maxi, p=1):
    a, b, mx = _prep(a, b, maxi)
    f = oracle.orc_gd_common
    f.restype = ctypes.c_double

    def gd(x, y, plus, psize, pp):
        return f(P(x), ctypes.c_size_t(x.shape[0]), P(y), ctypes.c_size_t(y.shape[0]), x.shape[1], mx.ctypes.data_as(u8p), plus, psize, pp)
    if what == "igd":
        return gd(b, a, 0, 0, 1)
    if what == "igd_plus":
        return gd(b, a, 1, 1, 1)
    return max(gd(a, b, 0, 1, p), gd(b, a, 0, 1, p))


def refc(L, what, a, b, maxi, p=1):
    a, b, mx = _prep(a, b, maxi)
    args = (P(a), ctypes.c_size_t(a.shape[0]), a.shape[1], P(b), ctypes.c_size_t(b.shape[0]), mx.ctypes.data_as(u8p))
    if what == "igd":
        L.ref_IGD.restype = ctypes.c_double
        return L.ref_IGD(*args)
    if what == "igd_plus":
        L.ref_IGD_plus.restype = ctypes.c_double
        return L.ref_IGD_plus(*args)
    L.ref_avg_Hausdorff_dist.restype = ctypes.c_double
    return L.ref_avg_Hausdorff_dist(*args, ctypes.c_uint(p))


# python/tests/test_moocore.py:174-222 and the doctest of python/src/moocore/_moocore.py:375-384
REF = np.array([10, 0, 6, 1, 2, 2, 1, 6, 0, 10], dtype=float).reshape((-1, 2))
A = np.array([4, 2, 3, 3, 2, 4], dtype=float).reshape((-1, 2))
B = np.array([8, 2, 4, 4, 2, 8], dtype=float).reshape((-1, 2))
DOC_DAT = np.array([[3.5, 5.5], [3.6, 4.1], [4.1, 3.2], [5.5, 1.5]])
DOC_REF = np.array([[1, 6], [2, 5], [3, 4], [4, 3], [5, 2], [6, 1]], dtype=float)
FLIP = np.array([[-1.0, 1.0]])
KATS = [
    ("igd", A, REF, False, 3.707092031609239), ("igd", B, REF, False, 2.591483465847630), ("igd", A, REF, True, 3.707092031609239),
    ("igd", B, REF, [True, False], 2.591483465847630), ("igd_plus", A, REF, False, 1.482842712474619),
    ("igd_plus", B, REF, False, 2.260112615949154), ("igd_plus", -A, -REF, True, 1.482842712474619),
    ("igd_plus", B * FLIP, REF * FLIP, [True, False], 2.260112615949154), ("hausdorff", A, REF, False, 3.707092031609239),
    ("hausdorff", B, REF, True, 2.591483465847630), ("igd", DOC_DAT, DOC_REF, False, 1.0627908666722465),
    ("igd_plus", DOC_DAT, DOC_REF, False, 0.9855036468106652), ("hausdorff", DOC_DAT, DOC_REF, False, 1.0627908666722465),
]


def random_cases():
    rng = np.random.default_rng(21)
    for i in range(30):
        d = int(rng.integers(2, 19))
        na, nb = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        a = rng.random((na, d))
        b = rng.random((nb, d))
        if i % 4 == 0:
            b[: min(na, nb)] = a[: min(na, nb)]              # zero distances
        maxi = rng.random(d) < 0.5 if i % 3 else bool(i % 2)
        yield a, b, maxi, int(rng.integers(1, 6))


@pytest.mark.parametrize("case", range(len(KATS)))
def test_oracle_known_answers(oracle, case):
    what, a, b, maxi, want = KATS[case]
    assert orc(oracle, what, a, b, maxi) == pytest.approx(want, rel=1e-14)


def test_oracle_port_equals_compiled_reference(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libepsref.so")
    if not os.path.exists(path):
        pytest.skip("compiled reference not present (needs /root/reference at build time)")
    L = ctypes.CDLL(path)
    for what, a, b, maxi, want in KATS:
        assert refc(L, what, a, b, maxi) == pytest.approx(want, rel=1e-14)
    for a, b, maxi, p in random_cases():
        for what in ("igd", "igd_plus", "hausdorff"):
            assert orc(oracle, what, a, b, maxi, p) == refc(L, what, a, b, maxi, p), (what, a.shape, b.shape, p)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(KATS)))
def test_gpu_known_answers(mb, case):
    what, a, b, maxi, want = KATS[case]
    fn = {"igd": mb.igd, "igd_plus": mb.igd_plus, "hausdorff": mb.avg_hausdorff_dist}[what]
    assert fn(a, b, maximise=maxi) == pytest.approx(want, rel=1e-14)


@pytest.mark.gpu
def test_gpu_bit_exact_vs_oracle(mb, oracle):
    for a, b, maxi, p in random_cases():
        assert mb.igd(a, b, maximise=maxi) == orc(oracle, "igd", a, b, maxi)
        assert mb.igd_plus(a, b, maximise=maxi) == orc(oracle, "igd_plus", a, b, maxi)
        assert mb.avg_hausdorff_dist(a, b, maximise=maxi, p=p) == orc(oracle, "hausdorff", a, b, maxi, p)
    rng = np.random.default_rng(2)
    a, b = rng.random((4000, 10)), rng.random((3001, 10))
    assert mb.igd(a, b) == orc(oracle, "igd", a, b, False)
    assert mb.igd_plus(a, b, maximise=[True] * 4 + [False] * 6) == orc(oracle, "igd_plus", a, b, [True] * 4 + [False] * 6)
    a, b = rng.random((60, 40)), rng.random((80, 40))
    assert mb.avg_hausdorff_dist(a, b, p=3) == orc(oracle, "hausdorff", a, b, False, 3)
