"""DZ2019-HW generator: accuracy against the oracle's long double Newton, and generator time.

Prints, per dimension, the worst and median relative difference of the generated directions to the oracle's
(windows at the start, the middle and the end of an N = 1e8 lattice), then the generator time of the cfg3 batch.
    gpurun -- 'python dev/hwgen_check.py'
"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
from moocore_b200 import _api

L = ctypes.CDLL(os.path.join(ROOT, "oracle", "libhvoracle.so"))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)

def odirs(d, N, i0, cnt):
    r = np.zeros((cnt, d))
    L.orc_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(r))
    return r

N = 100_000_000
for d in (2, 3, 4, 5, 7, 10, 13, 16, 20, 31):
    worst, med, wsum = 0.0, [], 0.0
    for i0 in (0, 1000, 12_345_678, 50_000_000, N - 3000):
        cnt = 3000
        got = _api.directions("DZ2019-HW", d, N, 0, i0, cnt)
        want = odirs(d, N, i0, cnt)
        # compare the weights w = 1/r (r = 1e20 marks w = 0): absolute difference, |w| <= 1
        wg, ww = 1.0 / got, 1.0 / want
        diff = np.abs(wg - ww)
        worst = max(worst, float(diff.max()))
        med.append(float(np.median(diff)))
        # the unit-norm check: sum of w^2 = 1
        wsum = max(wsum, float(np.max(np.abs((wg * wg).sum(axis=1) - 1.0)[:-1 if i0 + cnt >= N else None])))
    print("d=%2d  max |w - w_oracle| %.2e  median %.2e   max |sum w^2 - 1| %.2e" % (d, worst, float(np.median(med)), wsum))

for cfg in ("cfg3", "cfg2"):
    c = hvcases.CONFIGS[cfg]
    x, ref, maxi = hvcases.config_inputs(cfg)
    with mb.Plan(x, ref, maxi) as plan:
        for _ in range(2):
            hilo = plan.run(c["method"], c["nsamples"], 0, 0, c["nsamples"])
        st = plan.stats()
        print(cfg, "hv sum", repr(hilo), "maxmin %.3f ms dirgen %.3f ms newton capped %d" % (st["maxmin_ms"], st["dirgen_ms"], st["newton_capped"]))

# whole lattices over a small point set: generator time per dimension, and (with a -DHVB200_HW_CHECK build selected
# through HVB200_LIBRARY) the number of accepted single-evaluation solves that miss the stop rule
rng = np.random.default_rng(0)
for d in (3, 4, 6, 10, 16, 20, 28):
    Nd = 20_000_003 if d <= 20 else 4_000_037
    with mb.Plan(rng.random((300, d)), 1.1, allow_extension=True) as plan:
        plan.run("DZ2019-HW", Nd, 0, 0, Nd)
        st = plan.stats()
        print("d=%2d N=%d dirgen %.3f ms (%.3f ns/direction)  newton capped / check failures %d" % (d, Nd, st["dirgen_ms"], 1e6 * st["dirgen_ms"] / Nd, st["newton_capped"]))
