"""Scratch GPU check (run under gpurun): parity vs compiled reference + first timings."""
import ctypes, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moocore_b200 as m
from moocore_b200 import _api

R = ctypes.CDLL("oracle/_ref/libhvref.so")
dp = ctypes.POINTER(ctypes.c_double)
def P(a): return a.ctypes.data_as(dp)
for nm in ("hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus", "ref_expected_value", "ref_hw_window", "ref_mc_window", "ref_rphi_window"):
    getattr(R, nm).restype = ctypes.c_double
u8p = ctypes.POINTER(ctypes.c_uint8)

def ref_call(method, pts, ref, maxi, N, seed=0):
    pts = np.ascontiguousarray(pts, dtype=float); n, d = pts.shape
    ref = np.ascontiguousarray(np.broadcast_to(np.asarray(ref, float), (d,)))
    mx = np.ascontiguousarray(np.broadcast_to(np.asarray(maxi, bool), (d,))).view(np.uint8)
    a = (P(pts), ctypes.c_size_t(n), ctypes.c_uint8(d), P(ref), mx.ctypes.data_as(u8p), ctypes.c_uint64(N))
    if method == "DZ2019-MC": return R.hv_approx_normal(*a, ctypes.c_uint32(seed))
    if method == "DZ2019-HW": return R.hv_approx_hua_wang(*a)
    return R.hv_approx_rphi_fang_wang_plus(*a)

def front(kind, n, d, seed):
    rng = np.random.default_rng(seed)
    if kind == "sphere":
        x = np.abs(rng.normal(size=(n, d))); x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "simplex":
        x = rng.exponential(size=(n, d)); x /= x.sum(1, keepdims=True)
    else:
        x = rng.random((n, d))
    return x

print("devices", m.device_count())
print("fp64 bench", m.measure_fp64(0))

# 1. KAT
x = np.array([[5, 5], [4, 6], [2, 7], [7, 4]], float)
for meth in ("DZ2019-MC", "DZ2019-HW", "Rphi-FWE+"):
    g = m.hv_approx(x, ref=[10, 10], seed=42, method=meth); r = ref_call(meth, x, [10, 10], False, 262144, 42)
    print("KAT", meth, repr(g), repr(r), "rel", abs(g - r) / r)

# 2. directions
for d, N in ((3, 100000), (5, 1000000), (10, 100000000), (20, 1000000000)):
    for i0 in (0, N // 3, N - 4096):
        cnt = 4096
        r1 = np.zeros((cnt, d)); R.ref_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(r1))
        r2 = _api.directions("DZ2019-HW", d, N, 0, i0, cnt)
        rel = np.abs(r1 - r2) / np.abs(r1)
        print("HWdir d=%d N=%g i0=%d maxrel %.2e mean %.2e exact %.3f" % (d, N, i0, rel.max(), rel.mean(), (r1 == r2).mean()))

# 3. per-sample values bit-exact on reference directions
for kind, n, d in (("sphere", 100, 3), ("simplex", 1000, 5), ("sphere", 10000, 10), ("random", 5000, 20), ("sphere", 777, 7), ("random", 333, 31), ("random", 100, 2)):
    pts = 1.1 - front(kind, n, d, 1)
    N = 100000
    cnt = 3000
    r1 = np.zeros((cnt, d)); R.ref_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(1234), ctypes.c_uint64(cnt), P(r1))
    want = np.array([R.ref_expected_value(P(pts), ctypes.c_size_t(n), d, P(r1[i])) for i in range(cnt)])
    for kern in ("exact", "screen"):
        plan = m.Plan(1.1 - pts, 1.1, kernel=kern)   # Plan transforms: ref - p  => pass original coords
        vals, hilo = plan.run_directions(r1)
        st = plan.stats()
        print("values", kind, n, d, kern, "npoints", plan.npoints, "bit-exact:", bool((vals == want).all()), "maxrel %.2e" % (np.abs(vals - want) / want).max(),
              "sum rel %.2e" % (abs(hilo[0] + hilo[1] - want.sum()) / want.sum()), "slow", st["slow_path"])
        plan.close()

# 4. full estimates config 1, 2
pts = front("sphere", 100, 3, 1)
for meth in ("DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"):
    g = m.hv_approx(pts, ref=1.1, nsamples=100000, seed=42, method=meth); r = ref_call(meth, pts, 1.1, False, 100000, 42)
    print("cfg1", meth, repr(g), repr(r), "rel %.2e" % (abs(g - r) / r))
pts = front("simplex", 1000, 5, 2)
for meth in ("DZ2019-MC", "DZ2019-HW", "Rphi-FWE+"):
    t0 = time.time(); g = m.hv_approx(pts, ref=1.1, nsamples=1000000, seed=42, method=meth); t1 = time.time()
    r = ref_call(meth, pts, 1.1, False, 1000000, 42); t2 = time.time()
    print("cfg2", meth, repr(g), repr(r), "rel %.2e" % (abs(g - r) / r), "gpu %.3fs ref %.3fs" % (t1 - t0, t2 - t1))

# 5. config 3 timing on a slice
pts = front("sphere", 10000, 10, 3)
N = 100000000
for kern in ("exact", "screen"):
    plan = m.Plan(pts, 1.1, kernel=kern)
    for cnt in (1 << 20, 1 << 23):
        plan.run("DZ2019-HW", N, 0, 0, cnt)
        t0 = time.time(); hilo = plan.run("DZ2019-HW", N, 0, 0, cnt); t1 = time.time()
        st = plan.stats()
        sp = cnt * plan.npoints
        print("cfg3", kern, "cnt", cnt, "wall %.4fs maxmin %.2fms dirgen %.2fms  => maxmin %.1f Gsp/s  total %.1f Gsp/s  slow/sample %.1f capped %d" % (
            t1 - t0, st["maxmin_ms"], st["dirgen_ms"], sp / st["maxmin_ms"] / 1e6, sp / (t1 - t0) / 1e9, st["slow_path"] / cnt, st["newton_capped"]))
    # window parity vs reference
    cnt = 20000
    i0 = 50_000_000
    vals = plan.sample_values("DZ2019-HW", N, 0, i0, cnt)
    tp = np.ascontiguousarray(1.1 - pts)
    want = np.zeros(cnt)
    t0 = time.time(); s = R.ref_hw_window(P(tp), ctypes.c_size_t(10000), 10, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(want)); t1 = time.time()
    rel = np.abs(vals - want) / want
    print("cfg3 window", kern, "sum rel %.3e" % (abs(vals.sum() - want.sum()) / want.sum()), "max sample rel %.2e" % rel.max(), "exact frac %.3f" % (vals == want).mean(),
          "ref %.2f s => %.3f Gsp/s 1 core" % (t1 - t0, cnt * 10000 / (t1 - t0) / 1e9))
    plan.close()
