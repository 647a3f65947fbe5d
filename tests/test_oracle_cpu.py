"""CPU tests that PIN the oracle (oracle/hvapprox_oracle.c): against the reference's own
known-answer values, against golden vectors produced by the compiled reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the reference itself."""
import ctypes

import numpy as np
import pytest

import hvcases
from util import P, c_args, estimate, fh, rel

TOL = 1e-10   # BASELINE.json: relative tolerance on the estimate

KAT = np.array([[5, 5], [4, 6], [2, 7], [7, 4]], float)


@pytest.mark.parametrize("method", ["DZ2019-MC", "DZ2019-HW", "Rphi-FWE+"])
def test_kat_2d(oracle, golden, method):
    """python/src/moocore/_moocore.py:1093-1100 and r/tests/testthat/test-doctest-hv_approx.R:9-11."""
    got = estimate(oracle, "orc_", method, KAT, [10, 10], False, 262144, 42)
    assert got == fh(golden["kat2d"][method])                      # bit-exact with the compiled reference
    assert abs(got - golden["kat2d_doctest"][method]) < 1.5e-6     # the published doctest digits


def test_inputs_unchanged(golden):
    for name, dg in golden["digests"].items():
        x, _, _ = hvcases.config_inputs(name)
        assert hvcases.digest(x) == dg, f"numpy stream changed for {name}: regenerate goldens"


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
@pytest.mark.parametrize("method", ["DZ2019-MC", "DZ2019-HW", "Rphi-FWE+"])
def test_full_estimates(oracle, golden, name, method):
    c = hvcases.CONFIGS[name]
    x, ref, maxi = hvcases.config_inputs(name)
    got = estimate(oracle, "orc_", method, x, ref, maxi, c["nsamples"], 42)
    want = fh(golden[name][method])
    assert rel(got, want) <= TOL
    if method != "DZ2019-HW":
        assert got == want    # MC and Rphi restate the reference bit-for-bit
    # low-d cross-check against exact fpli_hv: tolerance = approximation error
    assert rel(got, fh(golden[name]["exact"])) < 5e-3


@pytest.mark.parametrize("name", ["cfg3", "cfg4"])
def test_hw_windows(oracle, golden, name):
    c = hvcases.CONFIGS[name]
    x, ref, maxi = hvcases.config_inputs(name)
    tp = hvcases.transformed(x, ref, maxi)
    wins = golden[name + "_windows"]
    if name == "cfg4":
        wins = wins[:4]       # ~1 s of oracle time per window
    for w in wins:
        cnt = w["count"] if name == "cfg3" else 64
        vals = np.zeros(cnt)
        s = oracle.orc_hw_window(P(tp), ctypes.c_size_t(tp.shape[0]), c["d"], ctypes.c_uint64(c["nsamples"]),
                                 ctypes.c_uint64(w["i0"]), ctypes.c_uint64(cnt), P(vals))
        want_vals = np.array([fh(v) for v in w["values"]])
        k = min(cnt, len(want_vals))
        # first windows have tiny u_0 -> ill-conditioned angles; per-sample tolerance is looser than the sum's
        assert np.max(np.abs(vals[:k] - want_vals[:k]) / want_vals[:k]) < 1e-7
        if cnt == w["count"]:
            assert rel(s, fh(w["sum"])) <= TOL


def test_hw_directions(oracle, golden):
    for e in golden["dirs"]:
        d, N, i0 = e["d"], e["N"], e["i0"]
        want = np.array([[fh(v) for v in row] for row in e["r"]])
        got = np.zeros_like(want)
        oracle.orc_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(want.shape[0]), P(got))
        tol = 1e-6 if i0 == 0 else 1e-9     # index 0: u_0 = 1/N, angle ~ u^(1/(m+1)) amplifies 1e-15
        assert np.max(np.abs(got - want) / np.abs(want)) < tol


def test_mc_stream(oracle, golden):
    z = np.zeros(4096)
    oracle.orc_mc_normals(ctypes.c_uint32(42), ctypes.c_uint64(4096), P(z))
    assert [float(v).hex() for v in z[:64]] == golden["mc_normals_seed42"]
    assert float(z.sum()).hex() == golden["mc_normals_seed42_sum4096"]


def test_transform_and_filter_edges(oracle):
    x = np.array([[1.0, 2.0, 3.0], [0.5, 0.5, 0.5], [2.0, 0.1, 0.1], [0.2, 0.2, 2.0]])
    ref = np.array([2.0, 2.0, 2.0])
    out = np.zeros_like(x)
    mx = np.zeros(3, np.uint8)
    n = oracle.orc_transform_and_filter(P(x), ctypes.c_size_t(4), 3, P(ref), mx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), P(out))
    assert n == 1 and np.array_equal(out[0], [1.5, 1.5, 1.5])     # equality with ref is NOT strict dominance
    mx = np.array([1, 0, 1], np.uint8)
    n = oracle.orc_transform_and_filter(P(x), ctypes.c_size_t(4), 3, P(np.array([0.1, 2.0, 0.4])), mx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), P(out))
    assert n == 2 and np.allclose(out[0], [0.4, 1.5, 0.1]) and np.allclose(out[1], [0.1, 1.8, 1.6])
    # no dominating point -> estimate 0 (python/tests/test_moocore.py:578)
    assert estimate(oracle, "orc_", "DZ2019-HW", x, [0.0, 0.0, 0.0], False, 1000) == 0.0
    assert estimate(oracle, "orc_", "DZ2019-MC", x, [0.0, 0.0, 0.0], False, 1000, 1) == 0.0
    assert estimate(oracle, "orc_", "Rphi-FWE+", x, [0.0, 0.0, 0.0], False, 1000) == 0.0


@pytest.mark.parametrize("dim", range(2, 12))
def test_reference_unit_test_shape(oracle, dim):
    """python/tests/test_moocore.py:532-554: single point 0.5, ref 1 -> exact HV 0.5**dim."""
    x = np.full((1, dim), 0.5)
    exact = 0.5 ** dim
    for method, sig in (("DZ2019-MC", 2), ("DZ2019-HW", 2), ("Rphi-FWE+", 2)):
        got = estimate(oracle, "orc_", method, x, 1.0, False, 262144 if dim < 9 else 65536, 42)
        assert abs(got - exact) / exact < 10.0 ** (-sig + 1)


# ---------------------------------------------------------------------------------------------
# directly against the compiled reference, when present
# ---------------------------------------------------------------------------------------------
def test_pow_chain_vs_reference(oracle, reflib):
    rng = np.random.default_rng(0)
    for e in range(0, 70):
        for x in rng.random(50) * 2:
            assert oracle.orc_pow_uint(x, e) == reflib.ref_pow_uint(x, e)


def test_int_sin_pow_vs_reference(oracle, reflib):
    rng = np.random.default_rng(1)
    for m in range(0, 19):
        for b in rng.random(200) * np.pi / 2:
            assert abs(oracle.orc_int_sin_pow(m, b) - reflib.ref_int_sin_pow(m, b)) < 5e-15


def test_zero_target_table_vs_reference(oracle, reflib):
    for m in range(0, 21):
        assert oracle.orc_solve_inverse_u(0.0, m) == reflib.ref_solve_inverse_u(0.0, m)


@pytest.mark.parametrize("d,N", [(3, 1000), (4, 77777), (7, 262144), (10, 100_000_000), (13, 5_000_000), (20, 1_000_000_000)])
def test_random_hw_windows_vs_reference(oracle, reflib, d, N):
    rng = np.random.default_rng(d)
    tp = np.ascontiguousarray(1.1 - hvcases.front("sphere", 200, d, d))
    for i0 in [0, int(rng.integers(0, N - 500)), N - 500]:
        cnt = min(500, N - i0)
        a = oracle.orc_hw_window(P(tp), ctypes.c_size_t(200), d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), None)
        b = reflib.ref_hw_window(P(tp), ctypes.c_size_t(200), d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), None)
        assert rel(a, b) <= TOL


@pytest.mark.parametrize("d", [2, 3, 6, 9, 16, 31])
def test_mc_rphi_windows_vs_reference_bit_exact(oracle, reflib, d):
    tp = np.ascontiguousarray(1.1 - hvcases.front("simplex", 100, d, d))
    v1, v2 = np.zeros(300), np.zeros(300)
    a = oracle.orc_mc_window(P(tp), ctypes.c_size_t(100), d, ctypes.c_uint32(7), ctypes.c_uint64(1000), ctypes.c_uint64(300), P(v1), None)
    b = reflib.ref_mc_window(P(tp), ctypes.c_size_t(100), d, ctypes.c_uint32(7), ctypes.c_uint64(1000), ctypes.c_uint64(300), P(v2), None)
    assert a == b and np.array_equal(v1, v2)
    a = oracle.orc_rphi_window(P(tp), ctypes.c_size_t(100), d, ctypes.c_uint64(1000), ctypes.c_uint64(300), P(v1), None, None)
    b = reflib.ref_rphi_window(P(tp), ctypes.c_size_t(100), d, ctypes.c_uint64(1000), ctypes.c_uint64(300), P(v2), None, None)
    assert a == b and np.array_equal(v1, v2)


# ---------------------------------------------------------------------------------------------
# round 2: config 5 (d = 50, patched reference), extension dimensions, DZ2019-HW at d = 21, 22
# ---------------------------------------------------------------------------------------------
def test_cfg5_inputs_unchanged(golden_r2):
    x, _, _ = hvcases.config_inputs("cfg5")
    assert hvcases.digest(x) == golden_r2["cfg5_digest"]


@pytest.mark.parametrize("w", [0, 1])
def test_cfg5_mc_windows(oracle, golden_r2, w):
    """Config 5 against the patched reference (oracle/patch_ref_d64.py): n = 1e5, d = 50, mixed maximise mask.
    The third window (1e7 samples = 5e8 normals deep) is left to the GPU test: 9 s of host RNG."""
    c = hvcases.CONFIGS["cfg5"]
    x, ref, maxi = hvcases.config_inputs("cfg5")
    tp = hvcases.transformed(x, ref, maxi)
    g = golden_r2["cfg5_windows"][w]
    assert tp.shape[0] == g["npoints"]
    cnt = 64
    vals, r = np.zeros(cnt), np.zeros((cnt, c["d"]))
    oracle.orc_mc_window(P(tp), ctypes.c_size_t(tp.shape[0]), c["d"], ctypes.c_uint32(c["seed"]),
                         ctypes.c_uint64(g["i0"]), ctypes.c_uint64(cnt), P(vals), P(r))
    assert [float(v).hex() for v in vals] == g["values"][:cnt]
    assert [[float(v).hex() for v in row] for row in r[:4]] == g["dirs"]


def test_cfg5_small_full_estimates(oracle, golden_r2):
    g = golden_r2["cfg5_small"]
    x, ref, maxi = hvcases.config_inputs("cfg5")
    xs = np.ascontiguousarray(x[:g["n"]])
    for method in ("DZ2019-MC", "Rphi-FWE+"):
        assert estimate(oracle, "orc_", method, xs, ref, maxi, g["nsamples"], 42) == fh(g[method])


def test_extension_dims_windows(oracle, golden_r2):
    for g in golden_r2["ext_dims"]:
        d = g["d"]
        tp = np.ascontiguousarray(1.1 - hvcases.front("random", g["n"], d, g["pseed"]))
        v = np.zeros(g["count"])
        s = oracle.orc_mc_window(P(tp), ctypes.c_size_t(g["n"]), d, ctypes.c_uint32(g["mc_seed"]), ctypes.c_uint64(g["i0"]),
                                 ctypes.c_uint64(g["count"]), P(v), None)
        assert s == fh(g["mc_sum"]) and [float(t).hex() for t in v[:32]] == g["mc_values"]
        s = oracle.orc_rphi_window(P(tp), ctypes.c_size_t(g["n"]), d, ctypes.c_uint64(g["i0"]), ctypes.c_uint64(g["count"]), P(v), None, None)
        assert s == fh(g["rphi_sum"]) and [float(t).hex() for t in v[:32]] == g["rphi_values"]


def test_hw_d21_d22_windows(oracle, golden_r2):
    """The last two dimensions where the reference's DZ2019-HW terminates."""
    for g in golden_r2["hw_d21_d22"]:
        d, N = g["d"], g["N"]
        tp = np.ascontiguousarray(1.1 - hvcases.front("sphere", g["n"], d, g["pseed"]))
        vals = np.zeros(g["count"])
        s = oracle.orc_hw_window(P(tp), ctypes.c_size_t(g["n"]), d, ctypes.c_uint64(N), ctypes.c_uint64(g["i0"]),
                                 ctypes.c_uint64(g["count"]), P(vals))
        assert rel(s, fh(g["sum"])) <= TOL
        want = np.array([fh(v) for v in g["values"]])
        assert np.max(np.abs(vals[:64] - want) / want) < (1e-6 if g["i0"] == 0 else 1e-9)


def test_int_sin_pow_vs_reference_m19_m20(oracle, reflib):
    """F_m at the two exponents only d = 21, 22 reach: the reduction formula and the reference's closed forms differ
    most here (both carry the cancellation noise that stops the reference from terminating at d >= 23)."""
    rng = np.random.default_rng(2)
    for m in (19, 20):
        for b in rng.random(2000) * np.pi / 2:
            assert abs(oracle.orc_int_sin_pow(m, b) - reflib.ref_int_sin_pow(m, b)) < 1e-14
        assert oracle.orc_solve_inverse_u(0.0, m) == reflib.ref_solve_inverse_u(0.0, m)


@pytest.mark.parametrize("d,N", [(21, 1_000_000), (22, 262144), (21, 100_000_000), (22, 1_000_000_000)])
def test_random_hw_windows_vs_reference_d21_d22(oracle, reflib, d, N):
    rng = np.random.default_rng(d)
    tp = np.ascontiguousarray(1.1 - hvcases.front("sphere", 200, d, d))
    for i0 in [0, int(rng.integers(0, N - 500)), N - 500]:
        a = oracle.orc_hw_window(P(tp), ctypes.c_size_t(200), d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(500), None)
        b = reflib.ref_hw_window(P(tp), ctypes.c_size_t(200), d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(500), None)
        assert rel(a, b) <= TOL


def test_patched_reference_is_the_reference_below_32(reflib):
    """The d <= 64 build differs from the reference only past its domain: same bits inside it."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libhvref_d64.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libhvref_d64.so not built")
    R64 = ctypes.CDLL(path)
    for nm in ("hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus"):
        getattr(R64, nm).restype = ctypes.c_double
    assert R64.ref_dimension_max() == 64 and reflib.ref_dimension_max() == 31
    for d in (2, 7, 20, 31):
        x = hvcases.front("sphere", 50, d, d)
        for method in ("DZ2019-MC", "Rphi-FWE+") + (("DZ2019-HW",) if d <= 20 else ()):
            assert estimate(R64, "", method, x, 1.1, False, 3000, 9) == estimate(reflib, "", method, x, 1.1, False, 3000, 9)
