"""Where does the pruned kernel start to pay INCLUDING its plan-time preparation (now on the device)?  Whole calls
(plan creation from host buffers + one DZ2019-HW run + result) with the DPX screen and with the pruned kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
shapes = [("sphere", 600, 5), ("sphere", 1000, 5), ("sphere", 2000, 10), ("random", 2000, 10), ("sphere", 5000, 10), ("sphere", 2000, 30),
          ("sphere", 20000, 10), ("random", 20000, 20), ("sphere", 100000, 3)]
for kind, n, d in shapes:
    x = hvcases.front(kind, n, d, 1)
    for N in (16384, 65536, 262144, 1048576):
        line = f"{kind:7s} n={n:6d} d={d:2d} N={N:8d}:"
        t = {}
        for kern in ("dpx", "pruned", "auto"):
            best = 1e9
            for rep in range(6):
                t0 = time.perf_counter()
                with mb.Plan(x, 1.1, kernel=kern) as plan:
                    hilo = plan.run("DZ2019-HW", N, 0)
                    used = plan.stats()["kernel"]
                best = min(best, time.perf_counter() - t0)
            t[kern] = best
            line += f"  {kern} {best * 1e3:7.3f} ms" + (f" (k={used})" if kern == "auto" else "")
        line += f"   pruned/dpx {t['pruned'] / t['dpx']:.2f}   samples/d = {N // d}"
        print(line, flush=True)
