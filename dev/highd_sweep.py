"""Pruned kernel at many objectives (d >= 24): phase C with and without the early-out vote between threshold planes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hvcases, moocore_b200 as mb
shapes = [("sphere", 2000, 30), ("random", 2000, 30), ("sphere", 20000, 24), ("random", 20000, 24), ("simplex", 20000, 31),
          ("sphere", 20000, 40), ("random", 20000, 40), ("random", 100000, 50), ("sphere", 100000, 50), ("random", 20000, 64)]
for kind, n, d in shapes:
    x = hvcases.front(kind, n, d, 1)
    meth = "DZ2019-HW" if d <= 31 else "DZ2019-MC"
    N = 50_000_000
    cnt = int(min(2e6, max(1e5, 2e10 / n)))
    i0 = N // 2 if meth == "DZ2019-HW" else 0
    line = f"{kind:8s} n={n:7d} d={d:2d} cnt={cnt:8d}:"
    res = []
    with mb.Plan(x, 1.1, kernel="pruned", allow_extension=d > 31) as plan:
        for flag in ("1", "0"):
            os.environ["HVB200_BP_EARLYOUT"] = flag
            for _ in range(2):
                hilo = plan.run(meth, N, 7, i0, i0 + cnt)
            st = plan.stats()
            res.append(hilo)
            line += f"  early_out={flag} {st['maxmin_ms']:9.3f} ms ({cnt * plan.npoints / st['maxmin_ms'] / 1e6:8.1f} Gsp/s)"
    line += "  same: %s" % (res[0] == res[1])
    print(line, flush=True)
os.environ.pop("HVB200_BP_EARLYOUT", None)
