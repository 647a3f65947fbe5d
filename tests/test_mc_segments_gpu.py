"""DZ2019-MC shards by pair ranges of the MT19937 stream (include/hvapprox_b200.h: hvb200_plan_mc_segment_scan,
hvb200_mc_segments_resolve, hvb200_plan_set_mc_start): the shards' per-sample values, concatenated, are the sequential
stream's values bit for bit (c/rng.c:71-99, c/hvapprox.c:235-254), whatever the cut points."""
import ctypes

import numpy as np
import pytest

import hvcases
import moocore_b200 as mb
from moocore_b200._api import MC_OPEN_END, mc_segment_layout, mc_segments_resolve

pytestmark = pytest.mark.gpu


def sharded_values(plan, N, seed, bounds):
    world = len(bounds)
    segs = [plan.mc_segment_scan(seed, bounds[g], (bounds[g + 1] - bounds[g]) if g + 1 < world else MC_OPEN_END)
            for g in range(world)]
    starts = mc_segments_resolve(segs, plan.nobj, N)
    parts = []
    for pair, discard, i0, i1 in starts:
        if i1 > i0:
            plan.set_mc_start(pair, discard)
            parts.append(plan.sample_values("DZ2019-MC", N, seed, i0, i1 - i0))
    return np.concatenate(parts), starts, segs


@pytest.mark.parametrize("d,seed", [(3, 42), (5, 0), (7, 2**32 - 1), (50, 7)])
def test_pair_range_shards_reproduce_the_sequential_stream(d, seed):
    n = 40
    x = hvcases.front("simplex", n, d, 3)
    N = 300_000 if d < 50 else 30_000
    rng = np.random.default_rng(d)
    with mb.Plan(x, np.full(d, 1.1), False, allow_extension=d > 31) as plan:
        want = plan.sample_values("DZ2019-MC", N, seed)
        total_pairs = int(N * d * 1.022)
        for world in (2, 3, 8):
            cuts = np.sort(rng.choice(np.arange(20000, total_pairs - 40000), size=world - 1, replace=False))
            bounds = [0, *[int(c) for c in cuts]]
            if min(np.diff(bounds), default=1 << 30) < 4 * 4096:
                bounds = [g * (total_pairs // world) for g in range(world)]
            got, starts, segs = sharded_values(plan, N, seed, bounds)
            assert got.shape == want.shape
            assert np.array_equal(got, want), (world, bounds, starts)
            # the skip state really varies between cuts (otherwise the test would not see a wrong s_in)
        # one run through plan.run as well: the sums of the shards add up to the sum of the whole
        segs = [plan.mc_segment_scan(seed, 0, total_pairs // 2), plan.mc_segment_scan(seed, total_pairs // 2, MC_OPEN_END)]
        starts = mc_segments_resolve(segs, d, N)
        tot = 0.0
        for pair, discard, i0, i1 in starts:
            plan.set_mc_start(pair, discard)
            hi, lo = plan.run("DZ2019-MC", N, seed, i0, i1)
            tot += hi + lo
        hi, lo = plan.run("DZ2019-MC", N, seed)
        assert abs(tot - (hi + lo)) <= 1e-12 * abs(hi + lo)


def test_layout_sized_segments_and_many_cut_points(monkeypatch):
    monkeypatch.setenv("HVB200_MC_SEGMENT_MIN", str(1 << 20))       # the library's default threshold is 2^26 pairs per rank
    # the library's own layout (ranges of >= 2^20 pairs), and cuts at every offset of a 40-pair window: all 16 incoming
    # skip states that occur in the stream are exercised
    d, n, seed = 4, 30, 42
    x = hvcases.front("sphere", n, d, 9)
    N = 1_200_000
    with mb.Plan(x, np.full(d, 1.1), False) as plan:
        want = plan.sample_values("DZ2019-MC", N, seed)
        lay = [mc_segment_layout(N, d, 4, g) for g in range(4)]
        assert all(l is not None for l in lay)
        got, starts, _ = sharded_values(plan, N, seed, [l[0] for l in lay])
        assert np.array_equal(got, want)
        # balance: the four sample ranges are within 1 % of each other
        sizes = [i1 - i0 for _, _, i0, i1 in starts]
        assert max(sizes) - min(sizes) < 0.01 * N / 4, sizes
        # cuts inside an attempt (skip_out > 0: ~1.3 % of the positions) and heads whose parses merge late (an attempt of
        # several pairs among the first 15: ~17 %) are the cases a wrong s_in or head count would break
        interesting, plain = [], []
        for off in range(600):
            cut = 1_000_000 + off
            a = plan.mc_segment_scan(seed, 0, cut)
            b = plan.mc_segment_scan(seed, cut, MC_OPEN_END)
            (interesting if a.skip_out or b.merge > 15 else plain).append(cut)
            if a.skip_out:
                assert len(set(b.head)) > 1 or b.merge > 15
        assert any(plan.mc_segment_scan(seed, 0, c).skip_out for c in interesting), "no cut inside an attempt among 600"
        for cut in interesting[:60] + plain[:10]:
            got, starts, segs = sharded_values(plan, N, seed, [0, cut])
            assert np.array_equal(got, want), (cut, starts)


def test_segment_scan_errors():
    d = 3
    x = hvcases.front("sphere", 20, d, 1)
    with mb.Plan(x, np.full(d, 1.1), False) as plan:
        with pytest.raises(RuntimeError):
            plan.mc_segment_scan(1, 0, 100)            # shorter than the head window


def test_set_mc_start_is_one_shot():
    d, N, seed = 3, 200_000, 5
    x = hvcases.front("sphere", 25, d, 2)
    with mb.Plan(x, np.full(d, 1.1), False) as plan:
        want = plan.sample_values("DZ2019-MC", N, seed, 0, 1000)
        segs = [plan.mc_segment_scan(seed, 0, 100_000), plan.mc_segment_scan(seed, 100_000, MC_OPEN_END)]
        pair, discard, i0, i1 = mc_segments_resolve(segs, d, N)[1]
        tail = plan.sample_values("DZ2019-MC", N, seed, i0, 500)
        plan.set_mc_start(pair, discard)
        assert np.array_equal(plan.sample_values("DZ2019-MC", N, seed, i0, 500), tail)        # used once ...
        assert np.array_equal(plan.sample_values("DZ2019-MC", N, seed, 0, 1000), want)        # ... and gone
        plan.set_mc_start(pair, discard)
        plan.run("DZ2019-HW", N, 0, 0, 1000)                                                  # any other run drops it
        assert np.array_equal(plan.sample_values("DZ2019-MC", N, seed, 0, 1000), want)
