"""Perf sweep across configs / kernels (slices of the full direction sets)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg1", "cfg2", "cfg3", "cfg4"]
kernels = [a for a in sys.argv[1:] if not a.startswith("cfg")] or ["exact", "screen", "dpx", "pruned"]
budget = {"cfg1": 100000, "cfg2": 1000000, "cfg3": 1 << 22, "cfg4": 1 << 19, "cfg5": 1 << 17}
for cfg in which:
    c = hvcases.CONFIGS[cfg]
    x, ref, maxi = hvcases.config_inputs(cfg)
    res = {}
    for kern in kernels:
        with mb.Plan(x, ref, maxi, kernel=kern, allow_extension=True) as plan:
            cnt = min(budget[cfg], c["nsamples"])
            meth = c["method"] if c["d"] <= 31 or c["method"] != "DZ2019-HW" else "DZ2019-MC"
            # DZ2019-HW: a window in the middle of the lattice (its first indices hug one axis and are unusually easy)
            i0 = (c["nsamples"] - cnt) // 2 if meth == "DZ2019-HW" else 0
            for _ in range(2):
                t0 = time.time(); hilo = plan.run(meth, c["nsamples"], c["seed"], i0, i0 + cnt); wall = time.time() - t0
            st = plan.stats()
            res[kern] = hilo[0] + hilo[1]
            print("%s %-6s d=%d n=%d cnt=%d  maxmin %.3f ms (%.1f Gsp/s)  dirgen %.3f ms  h2d %.3f ms  wall %.1f ms  slow/sample %.2f  capped %d" % (
                cfg, kern, c["d"], plan.npoints, cnt, st["maxmin_ms"], cnt * plan.npoints / max(st["maxmin_ms"], 1e-9) / 1e6,
                st["dirgen_ms"], st["h2d_ms"], wall * 1e3, st["slow_path"] / cnt, st["newton_capped"]), flush=True)
    print(cfg, "all sums equal:", len(set(res.values())) == 1, res)
