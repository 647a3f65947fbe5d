"""What dominated points cost the pruned kernel: cfg3's 10 000 nondominated points alone, and with 1x / 3x / 9x as
many dominated points mixed in (each a front point pulled towards the reference point).  The estimate must not
change by a bit; the question is the time.   gpurun -- 'python dev/dominated_cost.py'"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hvcases, moocore_b200 as mb

c = hvcases.CONFIGS["cfg3"]
x, ref, maxi = hvcases.config_inputs("cfg3")
ref = np.broadcast_to(np.asarray(ref, dtype=float), (x.shape[1],))
rng = np.random.default_rng(0)
N, cnt = c["nsamples"], 20_000_000
base = None
for mult in (0, 1, 3, 9):
    pts = x
    if mult:
        src = x[rng.integers(0, x.shape[0], mult * x.shape[0])]
        dom = ref - (ref - src) * rng.uniform(0.3, 0.98, (src.shape[0], 1))
        pts = np.vstack([x, dom])
        pts = pts[rng.permutation(pts.shape[0])]
    with mb.Plan(pts, ref, maxi) as plan:
        for _ in range(2):
            hilo = plan.run(c["method"], N, 0, 40_000_000, 40_000_000 + cnt)
        st = plan.stats()
    base = base or hilo
    print("n = %6d (%dx dominated)  max-min %.2f ms  %.2f G directions/s  same bits as the front alone: %s" % (
        pts.shape[0], mult, st["maxmin_ms"], cnt / st["maxmin_ms"] / 1e6, hilo == base))
