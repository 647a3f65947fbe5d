"""HypE weighted-hypervolume sampling (c/whv_hype.c): oracle port vs the compiled reference's golden values and the
reference's docstring examples (CPU); the ordered-sum closed form vs the literal loop (CPU, host code of the product
library); GPU entry points vs goldens and oracle, bit for bit (GPU)."""
import ctypes
import json
import os

import numpy as np
import pytest

import whvcases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_whv.json")))
fh = float.fromhex
DIST_ID = {"uniform": 0, "exponential": 1, "point": 2}


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def orc_whv(oracle, dist, pts, ideal, ref, nsamples, seed, mu_expo=None, goal=None):
    oracle.orc_whv_hype.restype = ctypes.c_double
    pts = np.ascontiguousarray(pts, dtype=float)
    ideal = np.ascontiguousarray(ideal, dtype=float)
    ref = np.ascontiguousarray(ref, dtype=float)
    mu = np.array([mu_expo, 0.0]) if dist == "exponential" else (np.array(goal, dtype=float) if dist == "point" else np.zeros(2))
    return oracle.orc_whv_hype(DIST_ID[dist], P(pts), pts.shape[0], P(ideal), P(ref), int(nsamples), ctypes.c_uint32(seed), P(mu))


# ---- CPU: the oracle is pinned to the compiled reference ------------------------------------------------------------
@pytest.mark.parametrize("name", list(whvcases.CASES))
def test_oracle_matches_reference_goldens(oracle, name):
    pts, ideal, ref, mu_expo, goal, nsamples, seed = whvcases.case_inputs(name)
    for dist in whvcases.DISTS:
        got = orc_whv(oracle, dist, pts, ideal, ref, nsamples, seed, mu_expo, goal)
        assert got == fh(GOLD["cases"][name][dist]), (name, dist)


def test_oracle_matches_reference_doctest(oracle):
    # python/src/moocore/_moocore.py:2832-2864
    for e in GOLD["doctest"]:
        got = orc_whv(oracle, e["dist"], np.array(e["points"], dtype=float), np.full(2, 1.0), np.full(2, 4.0), 100000, 42, e["mu"], e["mu"])
        assert got == fh(e["value"]) and abs(got - e["printed"]) < 1e-5


def test_oracle_matches_compiled_reference_when_present(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libwhvref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libwhvref.so not built (needs /root/reference)")
    R = ctypes.CDLL(path)
    for f in (R.whv_hype_unif, R.whv_hype_expo, R.whv_hype_gaus):
        f.restype = ctypes.c_double
    rng = np.random.default_rng(0)
    for _ in range(25):
        n, ns = int(rng.integers(0, 150)), int(rng.integers(1, 8000))
        pts = np.ascontiguousarray(rng.random((max(n, 1), 2)) * 3 + 1)[:n]
        ideal, ref = np.array([1.0, 1.0]) + 0.2 * rng.random(2), np.array([4.0, 4.0]) - 0.2 * rng.random(2)
        seed, mu, goal = int(rng.integers(0, 2**32 - 1)), float(0.05 + 0.5 * rng.random()), 2.0 + rng.random(2)
        args = (P(pts), n, P(ideal), P(ref), ns, ctypes.c_uint32(seed))
        assert R.whv_hype_unif(*args) == orc_whv(oracle, "uniform", pts, ideal, ref, ns, seed)
        assert R.whv_hype_expo(*args, ctypes.c_double(mu)) == orc_whv(oracle, "exponential", pts, ideal, ref, ns, seed, mu)
        assert R.whv_hype_gaus(*args, P(goal.copy())) == orc_whv(oracle, "point", pts, ideal, ref, ns, seed, goal=goal)


# ---- CPU: host code of the product library (no CUDA call) ------------------------------------------------------------
def test_ordered_sum_closed_form_equals_literal_loop(mb):
    L = mb.lib
    rng = np.random.default_rng(5)
    cases = [rng.integers(0, 40, 20000), rng.integers(0, 3, 50000), rng.integers(1, 5000, 3000), np.full(70000, 3),
             np.full(10000, 7), np.array([0, 0, 1, 0, 2]), np.zeros(10), rng.integers(0, 2, 300000) * rng.integers(1, 1000, 300000),
             np.full(40000, 1024), np.concatenate([np.full(5000, 3), np.full(5000, 4096), rng.integers(1, 100, 5000)])]
    for c in cases:
        c = np.ascontiguousarray(c, dtype=np.uint32)
        ptr = c.ctypes.data_as(ctypes.POINTER(ctypes.c_uint))
        fast = L.hvb200_whv_ordered_sum(ptr, c.size, 0)
        literal = L.hvb200_whv_ordered_sum(ptr, c.size, 1)
        assert fast == literal
        # and the literal loop is what it says
        if c.size <= 20000:
            s = 0.0
            for k in c.tolist():
                for _ in range(k):
                    s += 1.0 / k
            assert s == literal


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(whvcases.CASES))
def test_gpu_matches_goldens(mb, name):
    pts, ideal, ref, mu_expo, goal, nsamples, seed = whvcases.case_inputs(name)
    for dist in whvcases.DISTS:
        mu = {"uniform": None, "exponential": mu_expo, "point": goal}[dist]
        got = mb.whv_hype(pts, ref=ref, ideal=ideal, nsamples=nsamples, seed=seed, dist=dist, mu=mu)
        assert got == fh(GOLD["cases"][name][dist]), (name, dist)


@pytest.mark.gpu
def test_gpu_doctest_values(mb):
    for e in GOLD["doctest"]:
        got = mb.whv_hype(e["points"], ref=4, ideal=1, dist=e["dist"], mu=e["mu"], seed=42)
        assert got == fh(e["value"]) and abs(got - e["printed"]) < 1e-5


@pytest.mark.gpu
def test_gpu_matches_oracle_random(mb, oracle):
    rng = np.random.default_rng(77)
    for i in range(30):
        n, ns = int(rng.integers(1, 5000)), int(rng.integers(1, 60000))
        pts = 0.5 + 3.5 * rng.random((n, 2))
        if i % 4 == 0:
            pts[rng.integers(0, n, n // 3 + 1)] = pts[0]                 # duplicates
        ideal, ref = np.array([1.0, 1.0]) - 0.3 * rng.random(2), np.array([4.0, 4.0]) + 0.3 * rng.random(2)
        seed, mu_expo, goal = int(rng.integers(0, 2**32 - 1)), float(0.05 + 0.5 * rng.random()), 1.5 + 2.0 * rng.random(2)
        for dist in whvcases.DISTS:
            mu = {"uniform": None, "exponential": mu_expo, "point": goal}[dist]
            got = mb.whv_hype(pts, ref=ref, ideal=ideal, nsamples=ns, seed=seed, dist=dist, mu=mu)
            assert got == orc_whv(oracle, dist, pts, ideal, ref, ns, seed, mu_expo, goal), (i, dist)


@pytest.mark.gpu
def test_gpu_maximise_and_arguments(mb, oracle):
    rng = np.random.default_rng(3)
    pts = 1.0 + 3.0 * rng.random((200, 2))
    # maximising an objective = minimising its negation (_moocore.py:2878-2886)
    got = mb.whv_hype(pts, ref=[0.5, 4.5], ideal=[4.2, 0.8], maximise=[True, False], nsamples=20000, seed=5)
    want = orc_whv(oracle, "uniform", pts * np.array([-1.0, 1.0]), np.array([-4.2, 0.8]), np.array([-0.5, 4.5]), 20000, 5)
    assert got == want
    assert mb.whv_hype(pts, ref=4.5, ideal=0.8, nsamples=1000, seed=np.random.default_rng(1)) > 0
    with pytest.raises(NotImplementedError):
        mb.whv_hype(rng.random((5, 3)), ref=1, ideal=0)
    with pytest.raises(ValueError):
        mb.whv_hype(pts, ref=4.5, ideal=0.8, dist="nope")
    with pytest.raises(ValueError):
        mb.whv_hype(pts, ref=[1, 2, 3], ideal=0.8)
    # zero-width box: volume 0, and a point set that dominates nothing
    assert mb.whv_hype(pts, ref=[4.0, 1.0], ideal=[1.0, 1.0], nsamples=1000, seed=1) == 0.0
    assert mb.whv_hype(pts + 10.0, ref=4.5, ideal=0.8, nsamples=1000, seed=1) == 0.0


@pytest.mark.gpu
def test_gpu_large_counts_closed_form(mb, oracle):
    # every sample dominated by thousands of points: the ordered sum takes the closed form almost everywhere
    rng = np.random.default_rng(9)
    pts = 1.0 + 0.05 * rng.random((4000, 2))
    got = mb.whv_hype(pts, ref=4.0, ideal=1.0, nsamples=100000, seed=11)
    assert got == orc_whv(oracle, "uniform", pts, np.full(2, 1.0), np.full(2, 4.0), 100000, 11)
