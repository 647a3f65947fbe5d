"""Plan creation time with and without the nondominated pre-filter (worst case: a front, nothing to remove; and a
uniform cloud, where almost everything goes).   gpurun -- 'python dev/ndfilter_cost.py'"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import moocore_b200 as mb

rng = np.random.default_rng(0)
def sphere(n, d):
    v = np.abs(rng.standard_normal((n, d))) + 1e-3
    return v / np.linalg.norm(v, axis=1, keepdims=True)

for d in (3, 10, 30):
    for n in (2_000, 10_000, 32_768, 100_000, 262_144):
        if n * d > 2.2e7: continue
        for kind in ("front", "cloud"):
            tp = sphere(n, d) if kind == "front" else rng.random((n, d)) + 0.01
            res = []
            for flt in ("0", "1"):
                os.environ["HVB200_NDFILTER"] = flt
                best = 1e9
                for _ in range(3):
                    t = time.perf_counter()
                    plan = mb.Plan(tp, np.zeros(d), maximise=True)
                    best = min(best, time.perf_counter() - t)
                    kept = plan.npoints
                    plan.close() if hasattr(plan, "close") else plan.__exit__(None, None, None)
                res.append((best * 1e3, kept))
            print("d=%2d n=%8d %-5s  plan %.2f ms -> %.2f ms with the filter (kept %d)" % (d, n, kind, res[0][0], res[1][0], res[1][1]))
