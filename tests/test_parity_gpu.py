"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C-ABI of
libmoocore_b200.so; the CPU oracle (oracle/hvapprox_oracle.c), the committed goldens and — when
it travelled with the snapshot — the compiled reference (oracle/_ref) are the checkers.

Bars (BASELINE.json north_star):
  * per-sample values on a shared direction set: BIT-EXACT (max/min are exact, one rounding per
    product, the s^d chain has the reference's multiplication tree);
  * estimates / window sums with device-generated directions: <= 1e-10 relative;
  * low-d cross-check against exact fpli_hv with the approximation-error tolerance.
"""
import ctypes

import numpy as np
import pytest

import hvcases
from util import P, estimate, fh, rel

pytestmark = pytest.mark.gpu
TOL = 1e-10
METHODS = ["DZ2019-MC", "DZ2019-HW", "Rphi-FWE+"]
KAT = np.array([[5, 5], [4, 6], [2, 7], [7, 4]], float)


@pytest.fixture(scope="module", autouse=True)
def need_gpu(mb):
    assert mb.device_count() >= 1, "GPU tests need a CUDA device; there is no CPU fallback to test"


def oracle_values(oracle, tp, r):
    n, d = tp.shape
    return np.array([oracle.orc_expected_value(P(tp), ctypes.c_size_t(n), d, P(np.ascontiguousarray(r[i]))) for i in range(r.shape[0])])


def oracle_dirs(oracle, method, d, N, seed, i0, cnt):
    r = np.zeros((cnt, d))
    if method == "DZ2019-HW":
        oracle.orc_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(r))
    elif method == "DZ2019-MC":
        oracle.orc_mc_window(None, ctypes.c_size_t(0), d, ctypes.c_uint32(seed), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), None, P(r))
    else:
        oracle.orc_rphi_window(None, ctypes.c_size_t(0), d, ctypes.c_uint64(i0), ctypes.c_uint64(cnt), None, P(r), None)
    return r


# ---------------------------------------------------------------------------------------------
# known answers and full small configs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", METHODS)
def test_kat_2d(mb, golden, method):
    got = mb.hv_approx(KAT, ref=[10, 10], seed=42, method=method)
    assert rel(got, fh(golden["kat2d"][method])) <= TOL
    assert abs(got - golden["kat2d_doctest"][method]) < 1.5e-6


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
@pytest.mark.parametrize("method", METHODS)
def test_full_small_configs(mb, golden, name, method):
    c = hvcases.CONFIGS[name]
    x, ref, maxi = hvcases.config_inputs(name)
    got = mb.hv_approx(x, ref=ref, maximise=maxi, nsamples=c["nsamples"], seed=42, method=method)
    assert rel(got, fh(golden[name][method])) <= TOL
    assert rel(got, fh(golden[name]["exact"])) < 5e-3          # exact hv3d+/recursive cross-check


def test_cfg2_4d_slice_vs_exact_hv4d(mb, golden):
    c = hvcases.CONFIGS["cfg2"]
    x, ref, maxi = hvcases.config_inputs("cfg2")
    x4 = np.ascontiguousarray(x[:, :4])
    for method in ("DZ2019-MC", "DZ2019-HW"):
        got = mb.hv_approx(x4, ref=ref[:4], nsamples=c["nsamples"], seed=42, method=method)
        assert rel(got, fh(golden["cfg2"]["slice4_" + method])) <= TOL
        assert rel(got, fh(golden["cfg2"]["exact_4d_slice"])) < 1e-3


# ---------------------------------------------------------------------------------------------
# the max-min kernel on a shared direction set: bit-exact, both kernel variants, all dimensions
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", list(range(2, 33)) + [33, 40, 50, 64])
def test_values_bit_exact_all_dims(mb, oracle, d):
    n = 257 if d < 20 else 131
    x = hvcases.front("random" if d > 12 else "sphere", n, d, 100 + d)
    tp = np.ascontiguousarray(1.1 - x)
    method = "DZ2019-HW" if d <= 20 else "DZ2019-MC"
    r = oracle_dirs(oracle, method, d, 50_000, 5, 777, 600)
    want = oracle_values(oracle, tp, r)
    for kern in ("exact", "screen", "dpx", "pruned"):
        with mb.Plan(x, 1.1, kernel=kern, allow_extension=True) as plan:
            vals, hilo = plan.run_directions(r)
            assert np.array_equal(vals, want), f"d={d} kernel={kern}"
            assert rel(hilo[0] + hilo[1], float(np.sum(want))) < 1e-14


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 255, 256, 257, 1000, 4097, 30000])
def test_values_bit_exact_ragged_point_counts(mb, oracle, n):
    d = 6
    x = hvcases.front("simplex", n, d, n)
    tp = np.ascontiguousarray(1.1 - x)
    r = oracle_dirs(oracle, "DZ2019-MC", d, 10_000, 3, 0, 700 if n < 5000 else 300)
    want = oracle_values(oracle, tp, r)
    for kern in ("exact", "screen", "dpx", "pruned"):
        with mb.Plan(x, 1.1, kernel=kern) as plan:
            assert plan.npoints == n
            vals, _ = plan.run_directions(r)
            assert np.array_equal(vals, want), f"n={n} kernel={kern}"


def test_extreme_directions_bit_exact(mb, oracle):
    """Clamped weights (r = 1e20), huge/small magnitudes, ties between points."""
    d = 4
    x = np.array([[0.2, 0.3, 0.4, 0.5], [0.2, 0.3, 0.4, 0.5], [0.9, 0.1, 0.1, 0.1], [1e-300, 0.5, 0.5, 0.5], [0.6, 0.6, 0.6, 0.6]])
    r = np.array([[1e20, 1.0, 1.0, 1.0], [1.0, 1e20, 1e20, 1e20], [1.0, 1.0, 1.0, 1.0], [1e-10, 3.0, 7.0, 1e5],
                  [2.5, 2.5, 2.5, 2.5], [1e150, 1e150, 1e150, 1e150], [1e-150, 1.0, 1.0, 1.0]])
    tp = np.ascontiguousarray(1.0 - x)
    want = oracle_values(oracle, tp, r)
    for kern in ("exact", "screen", "dpx", "pruned"):
        with mb.Plan(x, 1.0, kernel=kern) as plan:
            vals, _ = plan.run_directions(r)
            assert np.array_equal(vals, want), kern


# ---------------------------------------------------------------------------------------------
# transform_and_filter through the boundary
# ---------------------------------------------------------------------------------------------
def test_no_dominating_point_returns_zero(mb):
    x = np.array([[1.0, 2.0], [3.0, 0.5]])
    for method in METHODS:
        assert mb.hv_approx(x, ref=[0.5, 0.5], method=method, seed=1) == 0.0     # python/tests/test_moocore.py:578
    assert mb.Plan(x, [0.5, 0.5]).empty


def test_filter_and_maximise_mask(mb, oracle):
    rng = np.random.default_rng(5)
    d = 5
    x = rng.random((400, d)) * 1.4                 # ~half of the rows do not dominate ref
    maxi = np.array([True, False, True, False, False])
    ref = np.where(maxi, 0.2, 1.1)
    tp = hvcases.transformed(x, ref, maxi)
    assert 0 < tp.shape[0] < 400
    with mb.Plan(x, ref, maxi) as plan:
        assert plan.npoints == tp.shape[0]
    for method in METHODS:
        got = mb.hv_approx(x, ref=ref, maximise=maxi, nsamples=20000, seed=9, method=method)
        want = estimate(oracle, "orc_", method, x, ref, maxi, 20000, 9)
        assert rel(got, want) <= TOL
    # maximise=True for all, like python/tests/test_moocore.py:557-566
    got = mb.hv_approx(-x, ref=-ref, maximise=~maxi, nsamples=20000, method="DZ2019-HW")
    want = estimate(oracle, "orc_", "DZ2019-HW", -x, -ref, ~maxi, 20000)
    assert rel(got, want) <= TOL


def test_dominated_points_do_not_change_the_result(mb):
    """python/src/moocore/_moocore.py:1104-1116: filter_dominated gives the same value."""
    x = hvcases.front("sphere", 300, 3, 4)
    extra = np.vstack([x, x[:50] + 0.05])          # dominated copies
    for method in ("DZ2019-HW", "Rphi-FWE+"):
        a = mb.hv_approx(x, ref=1.2, nsamples=30000, method=method)
        b = mb.hv_approx(extra, ref=1.2, nsamples=30000, method=method)
        assert a == b


# ---------------------------------------------------------------------------------------------
# direction generators
# ---------------------------------------------------------------------------------------------
def test_hw_directions_vs_golden(mb, golden):
    from moocore_b200 import _api
    for e in golden["dirs"]:
        d, N, i0 = e["d"], e["N"], e["i0"]
        want = np.array([[fh(v) for v in row] for row in e["r"]])
        got = _api.directions("DZ2019-HW", d, N, 0, i0, want.shape[0])
        tol = 1e-6 if (i0 == 0 or i0 > N - 100) else 1e-9
        assert np.max(np.abs(got - want) / np.abs(want)) < tol, (d, N, i0)


def test_hw_lattice_is_x87_exact(mb, oracle):
    """The lattice point u = frac(RN64((i+1)/N) * a_k) must reproduce the reference's long double
    arithmetic.  theta_0 has m = d-2; for d = 2, m = 0 and theta = u*pi/2 exactly,
    so r = (1/sin(theta), 1/cos(theta)) exposes u directly."""
    from moocore_b200 import _api
    for N in (1000, 99991, 2**20, 100_000_000, 2**31):
        idx = np.unique(np.concatenate([np.arange(0, 2000), N - 2000 + np.arange(0, 2000)]).clip(0, N - 1))
        for i0 in (0, N - 2000 if N > 4000 else 0):
            got = _api.directions("DZ2019-HW", 2, N, 0, i0, min(2000, N - i0))
            want = oracle_dirs(oracle, "DZ2019-HW", 2, N, 0, i0, min(2000, N - i0))
            assert np.max(np.abs(got - want) / np.abs(want)) < max(1e-12, 1e-15 * N)   # last samples: cos(theta) ~ 1/N amplifies 1e-16
            assert np.median(np.abs(got - want) / np.abs(want)) < 2e-14


def test_hw_zero_residue_samples(mb, oracle, golden):
    """Samples whose lattice coordinate has a zero residue (the reference's long double yields 0,
    a tiny positive number or ~1): windows around them must still match to 1e-10."""
    for name in ("cfg3",):
        c = hvcases.CONFIGS[name]
        x, ref, maxi = hvcases.config_inputs(name)
        with mb.Plan(x, ref, maxi) as plan:
            for w in golden[name + "_windows"][4:]:
                vals = plan.sample_values(c["method"], c["nsamples"], 0, w["i0"], w["count"])
                assert rel(float(np.sum(vals)), fh(w["sum"])) <= TOL
                want = np.array([fh(v) for v in w["values"]])
                assert np.max(np.abs(vals[:len(want)] - want) / want) < 1e-7


@pytest.mark.parametrize("name", ["cfg3", "cfg4"])
def test_hw_windows_vs_golden(mb, golden, name):
    c = hvcases.CONFIGS[name]
    x, ref, maxi = hvcases.config_inputs(name)
    with mb.Plan(x, ref, maxi) as plan:
        for w in golden[name + "_windows"]:
            vals = plan.sample_values(c["method"], c["nsamples"], 0, w["i0"], w["count"])
            assert rel(float(np.sum(vals)), fh(w["sum"])) <= TOL, (name, w["i0"])
        st = plan.stats()
        assert st["newton_capped"] == 0


@pytest.mark.parametrize("d", [2, 3, 5, 8, 10, 16, 31])
def test_mc_and_rphi_directions_match_oracle(mb, oracle, d):
    from moocore_b200 import _api
    for method, tol in (("DZ2019-MC", 0.0), ("Rphi-FWE+", 1e-13)):
        want = oracle_dirs(oracle, method, d, 100_000, 42, 5000, 1000)
        got = _api.directions(method, d, 100_000, 42, 5000, 1000)
        if tol == 0.0:
            assert np.array_equal(got, want), f"{method} d={d}"      # same stream, IEEE sqrt/div
        else:
            assert np.max(np.abs(got - want) / np.abs(want)) < tol   # CUDA pow/sincos vs glibc: ulps


# ---------------------------------------------------------------------------------------------
# sharding properties: range additivity, determinism, kernel equivalence at size
# ---------------------------------------------------------------------------------------------
def test_range_partials_add_up_and_are_deterministic(mb):
    c = hvcases.CONFIGS["cfg2"]
    x, ref, maxi = hvcases.config_inputs("cfg2")
    N = 200_000
    with mb.Plan(x, ref, maxi) as plan:
        for method in METHODS:
            full = plan.run(method, N, 42)
            again = plan.run(method, N, 42)
            assert full == again
            parts = []
            for g in range(4):
                parts.extend(plan.run(method, N, 42, N * g // 4, N * (g + 1) // 4))
            a = mb.hv_approx_finalize(c["d"], N, parts)
            b = mb.hv_approx_finalize(c["d"], N, full)
            assert rel(a, b) < 1e-14


def test_screen_equals_exact_at_bench_size_slice(mb):
    """cfg3 point set, 2^21 directions of the N=1e8 lattice: both kernels must give the same
    per-sample bits (size-independent property: the screen is exact)."""
    c = hvcases.CONFIGS["cfg3"]
    x, ref, maxi = hvcases.config_inputs("cfg3")
    cnt = 1 << 21
    out = {}
    for kern in ("exact", "screen", "dpx", "pruned"):
        with mb.Plan(x, ref, maxi, kernel=kern) as plan:
            out[kern] = plan.sample_values(c["method"], c["nsamples"], 0, 37_000_000, cnt)
    assert np.array_equal(out["exact"], out["screen"])
    assert np.array_equal(out["exact"], out["dpx"])
    assert np.array_equal(out["exact"], out["pruned"])
    assert np.all(out["exact"] > 0)


def test_cfg3_full_estimate(mb, golden):
    """The headline config at full size: 1e8 directions x 1e4 points.  Checked against the
    reference's own full run when the golden has it, and against the per-million window sums."""
    c = hvcases.CONFIGS["cfg3"]
    x, ref, maxi = hvcases.config_inputs("cfg3")
    got = mb.hv_approx(x, ref=ref, nsamples=c["nsamples"], method=c["method"])
    assert 0 < got < 1.1 ** 10
    if "cfg3_full" in golden:
        assert rel(got, fh(golden["cfg3_full"])) <= TOL
    if "cfg3_full_windows" in golden:
        with mb.Plan(x, ref, maxi) as plan:
            for w in (0, 1, 50, 99):
                hilo = plan.run(c["method"], c["nsamples"], 0, w * 1_000_000, (w + 1) * 1_000_000)
                assert rel(hilo[0] + hilo[1], fh(golden["cfg3_full_windows"][w])) <= TOL


def test_multi_gpu_single_process_matches_single(mb):
    if mb.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    x, ref, maxi = hvcases.config_inputs("cfg2")
    a = mb.hv_approx(x, ref=ref, nsamples=300_000, method="DZ2019-HW")
    import os
    os.environ["HVB200_NGPUS"] = "2"
    try:
        b = mb.hv_approx(x, ref=ref, nsamples=300_000, method="DZ2019-HW")
        from moocore_b200 import _api
        assert _api._load().hvb200_multi_last_exchange() == 1       # the partial sums went through ncclAllGather
        c = mb.hv_approx(x, ref=ref, nsamples=300_000, seed=42, method="DZ2019-MC")
        # the shards own pair ranges of the MT19937 stream (hvb200_mc_segment_layout; threshold lowered to this size) ...
        os.environ["HVB200_MC_SEGMENT_MIN"] = str(1 << 20)
        try:
            e = mb.hv_approx(x, ref=ref, nsamples=1_300_000, seed=42, method="DZ2019-MC")
        finally:
            del os.environ["HVB200_MC_SEGMENT_MIN"]
        os.environ["HVB200_MC_SHARD"] = "prefix"         # ... and the same estimate from contiguous sample ranges
        try:
            f = mb.hv_approx(x, ref=ref, nsamples=1_300_000, seed=42, method="DZ2019-MC")
        finally:
            del os.environ["HVB200_MC_SHARD"]
    finally:
        del os.environ["HVB200_NGPUS"]
    assert rel(a, b) < 1e-14
    assert rel(c, mb.hv_approx(x, ref=ref, nsamples=300_000, seed=42, method="DZ2019-MC")) < 1e-14
    single = mb.hv_approx(x, ref=ref, nsamples=1_300_000, seed=42, method="DZ2019-MC")
    assert rel(e, single) < 1e-14 and rel(f, single) < 1e-14


# ---------------------------------------------------------------------------------------------
# DZ2019-MC stream generated on the device (MT19937 + ziggurat, hvb200_mcgen.cuh)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 42, 2**32 - 1])
def test_mc_device_stream_bit_exact_across_batches(mb, oracle, seed):
    """Directions deep inside the stream (several internal batches, windows straddling the 2^20
    sample batch boundary) must equal the oracle's bit for bit: same MT19937 words, same ziggurat
    decisions, IEEE sqrt/div afterwards."""
    from moocore_b200 import _api
    d, N = 3, 2_500_000
    for i0 in (0, (1 << 20) - 500, 2_400_000):
        want = oracle_dirs(oracle, "DZ2019-MC", d, N, seed, i0, 1000)
        got = _api.directions("DZ2019-MC", d, N, seed, i0, 1000)
        assert np.array_equal(got, want), (seed, i0)


def test_mc_near_tie_decisions_go_to_the_host_generator(mb, monkeypatch):
    """an accept/reject comparison of the ziggurat's slow paths (c/rng.c:84-97) whose two sides agree to a few ulp
    could be decided differently by CUDA's and glibc's exp / log1p; such attempts are flagged and the run continues with the
    host generator.  With the band widened to 1e-3 the fallback really happens — and the values do not change."""
    x, ref, maxi = hvcases.config_inputs("cfg2")
    with mb.Plan(x, ref, maxi) as plan:
        dev = plan.sample_values("DZ2019-MC", 1_000_000, 9, 0, 400_000)
        monkeypatch.setenv("HVB200_MC_TIE_EPS", "1e-3")
        got = plan.sample_values("DZ2019-MC", 1_000_000, 9, 0, 400_000)
        st = plan.stats()
        assert st["h2d_bytes"] > 0, "the widened band must have sent the run to the host generator"
        monkeypatch.delenv("HVB200_MC_TIE_EPS")
        monkeypatch.setenv("HVB200_MC", "host")
        host = plan.sample_values("DZ2019-MC", 1_000_000, 9, 0, 400_000)
        assert np.array_equal(got, host)                 # the fallback IS the host stream
        # device vs host: the same accept/reject decisions everywhere (a different one would shift every later sample);
        # a tail value may differ in its last bits where CUDA's log1p and this box's glibc round differently
        bad = np.flatnonzero(dev != host)
        assert bad.size <= 40 and np.all(np.abs(dev[bad] - host[bad]) <= 1e-13 * np.abs(host[bad])), (bad.size, bad[:5])


def test_mc_device_equals_host_generator(mb, monkeypatch):
    x, ref, maxi = hvcases.config_inputs("cfg2")
    a = mb.hv_approx(x, ref=ref, nsamples=1_300_000, seed=7, method="DZ2019-MC")
    monkeypatch.setenv("HVB200_MC", "host")
    b = mb.hv_approx(x, ref=ref, nsamples=1_300_000, seed=7, method="DZ2019-MC")
    assert a == b


def test_mc_many_dimensions_stream(mb, oracle):
    """d = 31 consumes 31 normals per sample: checks the normal -> (sample, objective) mapping."""
    from moocore_b200 import _api
    want = oracle_dirs(oracle, "DZ2019-MC", 31, 200_000, 123, 150_000, 2000)
    got = _api.directions("DZ2019-MC", 31, 200_000, 123, 150_000, 2000)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------
# small-N and large-n edges
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 2, 3, 7, 33, 1000])
@pytest.mark.parametrize("d", [2, 3, 10])
def test_tiny_sample_counts(mb, oracle, N, d):
    """nsamples = 1 makes the only DZ2019-HW sample the all-zero 'last point' (c/hvapprox.c:300-303),
    whose angles are the reference's zero-target artefacts; small N also exercises partial chunks."""
    x = hvcases.front("sphere", 300, d, 50 + d)
    for method in METHODS:
        got = mb.hv_approx(x, ref=1.1, nsamples=N, seed=5, method=method)
        want = estimate(oracle, "orc_", method, x, 1.1, False, N, 5)
        assert rel(got, want) <= TOL, (method, N, d, got, want)


def test_large_point_set(mb, oracle):
    """2e5 points x 4 objectives: many tiles per chunk, keys and FP64 points beyond L1."""
    d, n = 4, 200_000
    x = hvcases.front("sphere", n, d, 77)
    tp = np.ascontiguousarray(1.1 - x)
    r = oracle_dirs(oracle, "DZ2019-HW", d, 100_000, 0, 40_000, 48)
    want = oracle_values(oracle, tp, r)
    for kern in ("exact", "screen", "dpx", "pruned"):
        with mb.Plan(x, 1.1, kernel=kern) as plan:
            vals, _ = plan.run_directions(r)
            assert np.array_equal(vals, want), kern


def test_plan_reuse_across_shapes(mb, oracle):
    """The workspace cache hands device buffers from one plan to the next: interleave shapes."""
    shapes = [(500, 3), (40, 12), (3000, 7), (10, 31), (500, 3)]
    for n, d in shapes:
        x = hvcases.front("random", n, d, n + d)
        got = mb.hv_approx(x, ref=1.2, nsamples=5000, method="DZ2019-HW")
        want = estimate(oracle, "orc_", "DZ2019-HW", x, 1.2, False, 5000)
        assert rel(got, want) <= TOL, (n, d)
