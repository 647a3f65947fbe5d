"""Latency of default-sized calls (nsamples = 262144, the Python/R default) through the drop-in entry points."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
for n, d in ((100, 3), (1000, 5), (10000, 10), (200, 20)):
    x = hvcases.front("sphere", n, d, 1)
    for method in ("Rphi-FWE+", "DZ2019-HW", "DZ2019-MC"):
        mb.hv_approx(x, ref=1.1, nsamples=262144, seed=1, method=method)
        t0 = time.perf_counter()
        for _ in range(20):
            mb.hv_approx(x, ref=1.1, nsamples=262144, seed=1, method=method)
        dt = (time.perf_counter() - t0) / 20
        print(f"n={n:6d} d={d:2d} {method:10s} {1e3 * dt:8.3f} ms per call  ({262144 * n / dt / 1e9:8.1f} Gsp/s)", flush=True)
