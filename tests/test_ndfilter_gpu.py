"""Nondominated pre-filter at plan creation (moocore_b200/csrc/hvb200_kernels.cu k_nd_flags).

A transformed point that another one weakly dominates can never attain max_p min_k p'_k r_k, so removing it must not
change one bit of any sample value or of the estimate.  Checked against a numpy brute-force filter (which points
survive, duplicates keep the first) and against runs with the filter switched off."""
import numpy as np
import pytest

import hvcases

pytestmark = pytest.mark.gpu


def brute_force_keep(tp):
    n = tp.shape[0]
    keep = np.ones(n, bool)
    for i in range(n):
        ge = (tp >= tp[i]).all(axis=1)
        gt = (tp > tp[i]).any(axis=1)
        dom = ge & (gt | (np.arange(n) < i))
        dom[i] = False
        keep[i] = not dom.any()
    return keep


def mixed_set(rng, n_front, n_dom, d, dup=0):
    front = 1.1 - hvcases.front("sphere", n_front, d, 3)              # transformed coordinates, all > 0
    src = front[rng.integers(0, n_front, n_dom)]
    dom = src * rng.uniform(0.3, 0.98, (n_dom, 1))
    parts = [front, dom]
    if dup:
        parts.append(front[rng.integers(0, n_front, dup)])           # exact duplicates of front points
    tp = np.vstack(parts)
    return np.ascontiguousarray(tp[rng.permutation(tp.shape[0])])


@pytest.mark.parametrize("d,n_front,n_dom,dup", [(2, 300, 900, 50), (3, 500, 700, 0), (6, 800, 1500, 100), (10, 1000, 3000, 10), (33, 400, 700, 20)])
@pytest.mark.parametrize("transform", ["host", "device"])
def test_filter_keeps_exactly_the_nondominated_points(mb, monkeypatch, d, n_front, n_dom, dup, transform):
    rng = np.random.default_rng(d)
    tp = mixed_set(rng, n_front, n_dom, d, dup)
    keep = brute_force_keep(tp)
    monkeypatch.setenv("HVB200_TRANSFORM", transform)
    out = {}
    for flt in ("0", "1"):
        monkeypatch.setenv("HVB200_NDFILTER", flt)
        # ref = 0 with maximise turns the points themselves into the transformed set
        with mb.Plan(tp, np.zeros(d), maximise=True, kernel="exact", allow_extension=True) as plan:
            out[flt] = (plan.npoints, plan.ndominated, plan.sample_values("DZ2019-MC", 3000, 5, 0, 3000),
                        plan.run("DZ2019-MC", 200_000, 5, 0, 200_000))
    assert out["0"][0] == tp.shape[0] and out["0"][1] == 0
    assert out["1"][0] == int(keep.sum()) and out["1"][1] == tp.shape[0] - int(keep.sum())
    assert np.array_equal(out["0"][2], out["1"][2])
    # the per-sample values carry the same bits; the double-double sum may group them differently (the EXACT
    # kernel sizes its grid by n'), which moves only the low word
    assert out["0"][3][0] == out["1"][3][0] and abs(out["0"][3][1] - out["1"][3][1]) <= 1e-15 * out["0"][3][0]
    # the survivors, in their original order, are the brute-force ones: a plan built from them gives the same values
    monkeypatch.setenv("HVB200_NDFILTER", "0")
    with mb.Plan(tp[keep], np.zeros(d), maximise=True, kernel="exact", allow_extension=True) as plan:
        assert np.array_equal(plan.sample_values("DZ2019-MC", 3000, 5, 0, 3000), out["1"][2])


def test_default_policy_and_entry_points(mb, monkeypatch):
    monkeypatch.delenv("HVB200_NDFILTER", raising=False)
    rng = np.random.default_rng(1)
    d = 5
    small = mixed_set(rng, 100, 300, d)                               # n' < 512: left alone
    with mb.Plan(small, np.zeros(d), maximise=True) as plan:
        assert plan.npoints == small.shape[0] and plan.ndominated == 0
    big = mixed_set(rng, 1500, 4500, d)
    with mb.Plan(big, np.zeros(d), maximise=True) as plan:
        assert plan.ndominated >= 4500 and plan.npoints + plan.ndominated == big.shape[0]
    # the drop-in entry points: same estimate with and without the filter, all three methods, all kernels
    x = 1.1 - big                                                     # minimisation input with ref = 1.1
    for method in ("DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"):
        got = {}
        for flt in ("0", "1"):
            monkeypatch.setenv("HVB200_NDFILTER", flt)
            got[flt] = mb.hv_approx(x, ref=1.1, nsamples=300_000, seed=3, method=method)
        assert got["0"] == got["1"], method
    for kern in ("dpx", "pruned"):
        vals = {}
        for flt in ("0", "1"):
            monkeypatch.setenv("HVB200_NDFILTER", flt)
            with mb.Plan(x, 1.1, kernel=kern) as plan:
                vals[flt] = plan.sample_values("DZ2019-HW", 1_000_003, 0, 400_000, 50_000)
        assert np.array_equal(vals["0"], vals["1"]), kern


def test_all_points_identical_or_one_dominating(mb, monkeypatch):
    monkeypatch.setenv("HVB200_NDFILTER", "1")
    d = 4
    tp = np.tile(np.array([0.3, 0.5, 0.7, 0.2]), (2000, 1))
    with mb.Plan(tp, np.zeros(d), maximise=True) as plan:
        assert plan.npoints == 1 and plan.ndominated == 1999
        a = plan.run("DZ2019-HW", 100_000, 0, 0, 100_000)
    tp[777] = [0.9, 0.9, 0.9, 0.9]                                    # dominates everything else
    with mb.Plan(tp, np.zeros(d), maximise=True) as plan:
        assert plan.npoints == 1 and plan.ndominated == 1999
        b = plan.run("DZ2019-HW", 100_000, 0, 0, 100_000)
    assert b[0] > a[0] > 0
