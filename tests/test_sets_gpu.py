"""hvb200_hv_approx_sets: many point sets against one direction set (the loop of c/main-hvapprox.c:149-174)
must return, per set, what the single-set entry points return."""
import time

import numpy as np
import pytest

import hvcases

pytestmark = pytest.mark.gpu
METHODS = ["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"]


@pytest.mark.parametrize("method", METHODS)
def test_sets_match_single_calls(mb, method):
    rng = np.random.default_rng(5)
    d = 4
    sets = [hvcases.front("sphere", int(n), d, 100 + i) for i, n in enumerate(rng.integers(1, 400, 25))]
    sets.append(np.full((3, d), 2.0))                     # nothing dominates ref -> 0.0
    sets.append(hvcases.front("simplex", 1, d, 7))        # a single point
    sets.append(np.zeros((0, d)))                         # an empty set
    got = mb.hv_approx_sets(sets, 1.1, nsamples=20000, seed=42, method=method)
    assert got.shape == (len(sets),)
    for s, g in zip(sets, got):
        want = mb.hv_approx(s, ref=1.1, nsamples=20000, seed=42, method=method) if s.shape[0] else 0.0
        assert g == pytest.approx(want, rel=1e-13, abs=0.0)
    assert got[-3] == 0.0 and got[-1] == 0.0 and got[-2] > 0


def test_sets_mixed_sizes_and_maximise(mb):
    # one set too large for the batched kernel (shared memory) goes through the single-set path
    d = 6
    maxi = np.array([True, False, False, True, False, False])
    rng = np.random.default_rng(9)
    sets = [rng.random((n, d)) for n in (50, 9000, 700, 3)]
    ref = np.where(maxi, -0.1, 1.1)
    got = mb.hv_approx_sets(sets, ref, maximise=maxi, nsamples=30000, method="DZ2019-HW")
    for s, g in zip(sets, got):
        assert g == pytest.approx(mb.hv_approx(s, ref=ref, maximise=maxi, nsamples=30000, method="DZ2019-HW"), rel=1e-13)


def test_sets_data_cumsizes_layout_and_high_d(mb):
    d = 20
    rng = np.random.default_rng(3)
    sizes = [10, 0, 120, 33]
    data = rng.random((sum(sizes), d))
    got = mb.hv_approx_sets((data, np.cumsum(sizes)), 1.05, nsamples=5000, seed=3, method="DZ2019-MC")
    lo = 0
    for n, g in zip(sizes, got):
        want = mb.hv_approx(data[lo:lo + n], ref=1.05, nsamples=5000, seed=3, method="DZ2019-MC") if n else 0.0
        assert g == pytest.approx(want, rel=1e-13, abs=0.0)
        lo += n


def test_sets_amortise_the_per_call_cost(mb):
    sets = [hvcases.front("sphere", 100, 3, i) for i in range(200)]
    mb.hv_approx_sets(sets, 1.1, nsamples=100_000, method="DZ2019-HW")           # warm-up: buffers of the full size
    mb.hv_approx(sets[0], ref=1.1, nsamples=100_000, method="DZ2019-HW")
    batched, looped = [], []
    for _ in range(3):               # best of three: one-off allocations and a noisy host must not decide the comparison
        t0 = time.perf_counter()
        a = mb.hv_approx_sets(sets, 1.1, nsamples=100_000, method="DZ2019-HW")
        t1 = time.perf_counter()
        b = [mb.hv_approx(s, ref=1.1, nsamples=100_000, method="DZ2019-HW") for s in sets]
        t2 = time.perf_counter()
        batched.append(t1 - t0)
        looped.append(t2 - t1)
    assert np.allclose(a, b, rtol=1e-13, atol=0)
    print("batched %.2f ms, looped %.2f ms" % (1e3 * min(batched), 1e3 * min(looped)))
    assert min(batched) < min(looped)


def test_rphi_direction_cache_is_transparent(mb, monkeypatch):
    # Rphi-FWE+ directions depend on (d, index) only and are kept on the device: growing runs, windows and other
    # dimensions in between must give the bits of the uncached streaming path
    x5 = hvcases.front("sphere", 300, 5, 1)
    x3 = hvcases.front("simplex", 300, 3, 2)
    calls = [(x5, 1000, 0, 1000), (x5, 400_000, 0, 400_000), (x3, 70_000, 10_000, 60_000), (x5, 400_000, 333_333, 399_999),
             (x5, 900_000, 250_000, 900_000), (x3, 300_000, 0, 300_000), (x5, 2000, 5, 1999)]
    got = []
    for x, N, i0, i1 in calls:
        with mb.Plan(x, 1.1) as plan:
            got.append(plan.run("Rphi-FWE+", N, 0, i0, i1))
    monkeypatch.setenv("HVB200_NO_DIRCACHE", "1")
    for (x, N, i0, i1), g in zip(calls, got):
        with mb.Plan(x, 1.1) as plan:
            assert plan.run("Rphi-FWE+", N, 0, i0, i1) == g
    r_cached = None
    monkeypatch.delenv("HVB200_NO_DIRCACHE")
    r_cached = mb.directions("Rphi-FWE+", 5, 500_000, 0, 123_456, 1000)
    monkeypatch.setenv("HVB200_NO_DIRCACHE", "1")
    assert np.array_equal(r_cached, mb.directions("Rphi-FWE+", 5, 500_000, 0, 123_456, 1000))
