"""Hypervolume-contribution approximation (second half): moocore_b200.hv_contributions_approx.

The reference has no approximate implementation (c/hv_contrib.c is exact), so parity is UNPINNED; the checks are
  * against the exact hv_contributions of the compiled reference (tests/golden/golden_contrib.json, both
    ignore_dominated semantics), with the tolerance of the estimator's own error (calibrated: 1e-4 of the largest
    contribution at 2^18 directions);
  * against a brute-force restatement (numpy top-2 on the directions the device generated): same arithmetic, 1e-12;
  * internal consistency: the hypervolume of the same pass equals hv_approx, two calls give the same bits."""
import json
import os

import numpy as np
import pytest

import hvcases
from util import fh, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cases():
    """the inputs of tests/golden/make_golden_contrib.py"""
    out = {}
    for name, kind, n, d, seed in (("sphere3", "sphere", 60, 3, 31), ("simplex4", "simplex", 40, 4, 32), ("sphere2", "sphere", 25, 2, 33)):
        x = hvcases.front(kind, n, d, seed)
        extra = np.vstack([x[:5] + 0.03, x[7:8], np.full((1, d), 1.3)])
        out[name] = (np.ascontiguousarray(np.vstack([x, extra])), np.full(d, 1.1))
    return out


@pytest.fixture(scope="module")
def golden_contrib():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_contrib.json")))


@pytest.mark.parametrize("name", ["sphere2", "sphere3", "simplex4"])
@pytest.mark.parametrize("ignore_dominated", [True, False])
def test_vs_exact_contributions(mb, golden_contrib, name, ignore_dominated):
    x, ref = cases()[name]
    g = golden_contrib[name]
    assert hvcases.digest(x) == g["digest"]
    want = np.array([fh(v) for v in g["contributions" if ignore_dominated else "contributions_keep_dominated"]])
    got, hv = mb.hv_contributions_approx(x, ref=ref, nsamples=1 << 20, method="DZ2019-HW", ignore_dominated=ignore_dominated, return_hv=True)
    assert np.array_equal(got == 0.0, want == 0.0)           # dominated rows, the duplicate pair, the row outside ref
    assert np.max(np.abs(got - want)) <= 2e-3 * want.max()
    assert rel(got.sum(), want.sum()) <= 1e-3
    assert rel(hv, fh(g["hv"])) <= 1e-3


@pytest.mark.parametrize("method", ["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"])
@pytest.mark.parametrize("d,n", [(3, 300), (10, 700), (13, 257), (31, 90)])
def test_vs_bruteforce_on_device_directions(mb, method, d, n):
    from math import gamma, pi
    from moocore_b200 import _api
    if method == "DZ2019-HW" and d > 20:
        pytest.skip("the reference's DZ2019-HW stops near d = 22")
    N = 20_000
    x = hvcases.front("sphere" if d <= 12 else "random", n, d, 200 + d)
    r = _api.directions(method, d, N, 5, 0, N)
    t = 1.1 - x
    v = np.empty((N, n))
    for a in range(0, N, 2000):
        v[a:a + 2000] = (t[None, :, :] * r[a:a + 2000, None, :]).min(axis=2)
    order = np.argsort(-v, axis=1, kind="stable")
    rows = np.arange(N)
    s1, s2, arg = v[rows, order[:, 0]], v[rows, order[:, 1]], order[:, 0]

    def pw(s):          # x^d without libm pow: repeated multiplication keeps the comparison at the 1e-12 level
        out = np.ones_like(s)
        for _ in range(d):
            out = out * s
        return out

    want = np.zeros(n)
    np.add.at(want, arg, pw(s1) - pw(s2))
    cd = 2 * pi ** (d / 2) / (gamma(d / 2) * d * 2 ** d)
    want *= cd / N
    got, hv = mb.hv_contributions_approx(x, ref=1.1, nsamples=N, seed=5, method=method, ignore_dominated=False, return_hv=True)
    scale = want.max()
    assert np.max(np.abs(got - want)) <= 1e-11 * scale
    assert rel(hv, mb.hv_approx(x, ref=1.1, nsamples=N, seed=5, method=method)) <= 1e-12


def test_deterministic_and_edge_cases(mb):
    x = hvcases.front("sphere", 1500, 5, 77)
    a = mb.hv_contributions_approx(x, ref=1.1, nsamples=300_000, method="DZ2019-HW")
    b = mb.hv_contributions_approx(x, ref=1.1, nsamples=300_000, method="DZ2019-HW")
    assert np.array_equal(a, b)                               # sorted accumulation: no atomics, same bits
    assert (a >= 0).all() and a.sum() > 0
    # nothing dominates ref -> zeros; a single point -> its contribution is the whole hypervolume
    assert not mb.hv_contributions_approx(np.array([[1.0, 2.0], [3.0, 0.5]]), ref=[0.5, 0.5], nsamples=1000).any()
    one, hv = mb.hv_contributions_approx(np.full((1, 4), 0.5), ref=1.0, nsamples=100_000, method="DZ2019-HW", return_hv=True)
    assert rel(one[0], hv) <= 1e-12 and rel(hv, 0.5 ** 4) < 1e-3
    with pytest.raises((ValueError, RuntimeError)):
        mb.hv_contributions_approx(np.full((2, 40), 0.5), ref=1.0, nsamples=100)      # more than 31 objectives


def test_maximise_mask(mb):
    rng = np.random.default_rng(3)
    x = rng.random((200, 4))
    maxi = np.array([True, False, False, True])
    ref = np.where(maxi, -0.1, 1.1)
    a = mb.hv_contributions_approx(x, ref=ref, maximise=maxi, nsamples=100_000, method="DZ2019-HW")
    y = np.where(maxi[None, :], 1.0 - x, x)                    # the same set, all-minimise
    b = mb.hv_contributions_approx(y, ref=1.1, nsamples=100_000, method="DZ2019-HW")
    assert np.max(np.abs(a - b)) <= 1e-12 * max(a.max(), 1e-300)
