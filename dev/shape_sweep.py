"""DPX vs PRUNED on odd shapes (mid-lattice DZ2019-HW slices), to check the AUTO kernel choice."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
shapes = [("sphere", 1_000_000, 3), ("sphere", 200_000, 5), ("sphere", 2000, 30), ("simplex", 512, 8), ("random", 700, 4),
          ("sphere", 100_000, 10), ("simplex", 20_000, 16), ("sphere", 5000, 2)]
N = 50_000_000
for kind, n, d in shapes:
    x = hvcases.front(kind, n, d, 1)
    budget = int(min(4e6, max(2e5, 2e11 / n)))
    line = f"{kind:8s} n={n:8d} d={d:2d} cnt={budget:8d}:"
    res = {}
    for kern in ("dpx", "pruned"):
        with mb.Plan(x, 1.1, kernel=kern) as plan:
            for _ in range(2):
                hilo = plan.run("DZ2019-HW", N, 0, N // 2, N // 2 + budget)
            st = plan.stats()
            res[kern] = hilo
            line += f"  {kern} {st['maxmin_ms']:9.3f} ms ({budget * plan.npoints / st['maxmin_ms'] / 1e6:9.1f} Gsp/s)"
    line += "  same hi: %s" % (res["dpx"][0] == res["pruned"][0] or abs(res["dpx"][0] - res["pruned"][0]) < 1e-12 * res["dpx"][0])
    print(line, flush=True)
