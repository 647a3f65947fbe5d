"""Adversarial inputs for the block-pruned max-min kernel (moocore_b200/csrc/hvb200_bp.cuh).

The kernel discards blocks and points through conservative 16-bit key tests and evaluates only the
survivors exactly, so every per-sample value must stay bit-identical to the textbook loop
(get_expected_value, c/hvapprox.c:68-117) whatever the point set looks like: ties, duplicates,
dominated points, constant objectives, extreme dynamic ranges, clusters below the key resolution,
candidate lists that overflow.  Checker: the CPU oracle on a subsample, the EXACT kernel on everything."""
import ctypes

import numpy as np
import pytest

import hvcases
from test_parity_gpu import oracle_dirs, oracle_values

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def no_nondominated_prefilter(monkeypatch):
    # these sets are built from duplicates and dominated points on purpose: keep them in front of the kernel
    monkeypatch.setenv("HVB200_NDFILTER", "0")


def check(mb, oracle, tp, r, n_oracle=300):
    """tp: transformed points (all > 0), r: directions.  exact == pruned everywhere, == oracle on a subsample."""
    tp = np.ascontiguousarray(tp, dtype=float)
    d = tp.shape[1]
    out = {}
    for kern in ("exact", "pruned"):
        # ref = 0 with maximise turns the points themselves into the transformed set
        with mb.Plan(tp, np.zeros(d), maximise=True, kernel=kern, allow_extension=True) as plan:
            assert plan.npoints == tp.shape[0]
            vals, hilo = plan.run_directions(r)
            out[kern] = vals
            assert plan.stats()["kernel"] == {"exact": 1, "pruned": 4}[kern]
    assert np.array_equal(out["exact"], out["pruned"])
    idx = np.linspace(0, r.shape[0] - 1, min(n_oracle, r.shape[0])).astype(int)
    want = oracle_values(oracle, tp, r[idx])
    assert np.array_equal(out["pruned"][idx], want)


def hw_dirs(oracle, d, cnt, N=1_000_003, i0=300_000):
    return oracle_dirs(oracle, "DZ2019-HW" if d <= 20 else "DZ2019-MC", d, N, 11, i0, cnt)


def test_identical_points_overflow_the_hit_list(mb, oracle):
    # every point ties with the maximum: every (direction, block) pair holds 32 candidates, far more
    # entries than the per-CTA hit list takes (kBpHitCap) -> the in-place evaluation path runs
    d, n = 6, 4096
    p = np.random.default_rng(0).random(d) + 0.2
    tp = np.tile(p, (n, 1))
    check(mb, oracle, tp, hw_dirs(oracle, d, 3000))


def test_duplicates_and_dominated_points(mb, oracle):
    rng = np.random.default_rng(1)
    d = 7
    front = 1.1 - hvcases.front("sphere", 1500, d, 9)
    dominated = front[rng.integers(0, 1500, 3000)] * rng.uniform(0.2, 1.0, (3000, 1))
    tp = np.vstack([front, front[:700], dominated])
    tp = tp[rng.permutation(tp.shape[0])]
    check(mb, oracle, tp, hw_dirs(oracle, d, 4000))


def test_constant_objective_and_two_level_columns(mb, oracle):
    rng = np.random.default_rng(2)
    d, n = 5, 3000
    tp = rng.random((n, d)) + 0.05
    tp[:, 1] = 0.75                                     # zero range -> key scale 0
    tp[:, 3] = np.where(rng.random(n) < 0.5, 0.3, 0.9)  # two values only
    check(mb, oracle, tp, hw_dirs(oracle, d, 4000))


def test_extreme_dynamic_range(mb, oracle):
    rng = np.random.default_rng(3)
    d, n = 8, 5000
    tp = 10.0 ** rng.uniform(-12, 6, (n, d))
    tp[:50] = 10.0 ** rng.uniform(-300, 300, (50, d))
    r = hw_dirs(oracle, d, 3000)
    r[::7] *= 1e150
    r[3::11] *= 1e-150
    check(mb, oracle, tp, r)


def test_cluster_below_the_key_resolution(mb, oracle):
    # all points within 1e-9 of one another: every quantised key coincides, only the exact path separates them
    rng = np.random.default_rng(4)
    d, n = 10, 2500
    base = rng.random(d) + 0.5
    tp = base * (1.0 + 1e-9 * rng.random((n, d)))
    tp[0] = base * 3.0                                   # one far outlier stretches the key range
    check(mb, oracle, tp, hw_dirs(oracle, d, 3000))


@pytest.mark.parametrize("d,n", [(2, 600), (3, 33), (4, 1), (9, 31), (17, 2000), (33, 700), (64, 520)])
def test_small_and_odd_shapes(mb, oracle, d, n):
    rng = np.random.default_rng(100 + d)
    tp = rng.random((n, d)) + 1e-3
    check(mb, oracle, tp, hw_dirs(oracle, d, 1500), n_oracle=150)


def test_many_directions_all_kernels_agree(mb):
    c = hvcases.CONFIGS["cfg4"]
    x, ref, maxi = hvcases.config_inputs("cfg4")
    x = x[:20000]
    out = {}
    for kern in ("dpx", "pruned"):
        with mb.Plan(x, ref, maxi, kernel=kern) as plan:
            out[kern] = plan.sample_values(c["method"], c["nsamples"], 0, 123_456_789, 300_000)
    assert np.array_equal(out["dpx"], out["pruned"])


def test_auto_kernel_choice(mb):
    x, ref, maxi = hvcases.config_inputs("cfg2")
    with mb.Plan(x, ref, maxi) as plan:                       # n = 1000
        plan.run("DZ2019-HW", 100_000, 0, 0, 100_000)
        assert plan.stats()["kernel"] == 3                    # 1e8 sample*points: not worth the preparation
        plan.run("DZ2019-HW", 10_000_000, 0, 0, 8_000_000)
        assert plan.stats()["kernel"] == 4
        plan.run("DZ2019-HW", 100_000, 0, 0, 100_000)
        assert plan.stats()["kernel"] == 4                    # prepared once, kept
    with mb.Plan(x[:200], ref, maxi) as plan:
        plan.run("DZ2019-HW", 10_000_000, 0, 0, 8_000_000)
        assert plan.stats()["kernel"] == 1


def test_block_cyclic_shards_sum_to_the_full_run(mb):
    # hvb200_plan_run_blocks: the interleaved shards of a DZ2019-HW lattice partition it exactly
    x, ref, maxi = hvcases.config_inputs("cfg2")
    N = 3_000_017
    with mb.Plan(x, ref, maxi) as plan:
        full = sum(plan.run("DZ2019-HW", N, 0, 0, N))
        for world, block in ((4, 1 << 18), (3, 100_000), (8, 1 << 21)):
            parts = [plan.run_blocks("DZ2019-HW", N, g, world, block=block) for g in range(world)]
            assert abs(sum(hi + lo for hi, lo in parts) - full) <= 1e-13 * full
            assert all(hi > 0 for hi, lo in parts) or block * (world - 1) >= N
        with pytest.raises(RuntimeError):
            plan.run_blocks("DZ2019-MC", N, 0, 2)
        with pytest.raises(RuntimeError):
            plan.run_blocks("DZ2019-HW", N, 2, 2)


def test_results_do_not_depend_on_the_schedule(mb):
    # per-chunk partial sums with a fixed tree: repeated runs (dynamic chunk hand-out) give identical bits
    x, ref, maxi = hvcases.config_inputs("cfg2")
    with mb.Plan(x, ref, maxi, kernel="pruned") as plan:
        runs = {plan.run("DZ2019-HW", 5_000_000, 0, 1_000_000, 4_000_000) for _ in range(5)}
    assert len(runs) == 1


@pytest.mark.parametrize("d,n", [(3, 5000), (10, 40000), (40, 9000)])
def test_device_transform_and_filter_matches_host(mb, d, n, monkeypatch):
    # transform_and_filter (c/hvapprox.c:30-66) on the device vs the host loop: same survivors, same order, same bits
    rng = np.random.default_rng(d)
    x = rng.random((n, d))
    maxi = rng.random(d) < 0.4
    ref = np.where(maxi, 0.25, 0.8)                  # about half of the coordinates fail -> most points are filtered
    x[:, :2] = np.where(maxi[:2], 0.25 + 0.7 * x[:, :2], 0.8 * x[:, :2])    # keep a few thousand points alive
    x[::2, 2:] = np.where(maxi[2:], 0.3 + 0.6 * x[::2, 2:], 0.7 * x[::2, 2:])
    x[5] = ref                                       # exactly on the reference point: dropped (not > 0)
    out = {}
    for mode in ("host", "device"):
        monkeypatch.setenv("HVB200_TRANSFORM", mode)
        with mb.Plan(x, ref, maxi, kernel="exact", allow_extension=True) as plan:
            out[mode] = (plan.npoints, plan.sample_values("DZ2019-MC", 4000, 7, 0, 4000))
    want = hvcases.transformed(x, ref, maxi).shape[0]
    assert out["host"][0] == out["device"][0] == want and 0 < want < n
    assert np.array_equal(out["host"][1], out["device"][1])
    # nothing survives: the entry point returns 0 like the reference (c/hvapprox.c:228-229)
    monkeypatch.setenv("HVB200_TRANSFORM", "device")
    if d <= 31:
        assert mb.hv_approx(x, ref=np.where(maxi, 2.0, -1.0), maximise=maxi, nsamples=1000, method="DZ2019-MC", seed=1) == 0.0
    with mb.Plan(x, np.where(maxi, 2.0, -1.0), maxi, allow_extension=True) as plan:
        assert plan.empty and plan.npoints == 0


def test_workspace_reuse_across_grid_sizes(mb):
    # plans hand their device buffers to a per-device pool; a d=20 plan (2 CTAs/SM) followed by a d=5 plan
    # (3 CTAs/SM, more hit lists) followed by d=20 again must not see stale per-chunk buffers
    rng = np.random.default_rng(8)
    for d, n in ((20, 3000), (5, 3000), (20, 2500), (6, 4000), (33, 1500)):
        x = rng.random((n, d))
        out = {}
        for kern in ("exact", "pruned"):
            with mb.Plan(x, 1.1, kernel=kern, allow_extension=True) as plan:
                out[kern] = plan.run("DZ2019-MC", 600_000, 9, 0, 300_000)
        assert abs(sum(out["exact"]) - sum(out["pruned"])) <= 1e-12 * sum(out["exact"])
