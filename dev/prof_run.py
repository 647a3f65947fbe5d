"""Short profiling target: cfg3 point set, 2^21 directions of the N=1e8 lattice, one kernel variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hvcases, moocore_b200 as mb
kern = sys.argv[1] if len(sys.argv) > 1 else "screen"
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 21)
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
i0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
c = hvcases.CONFIGS[cfg]
x, ref, maxi = hvcases.config_inputs(cfg)
with mb.Plan(x, ref, maxi, kernel=kern, allow_extension=True) as plan:
    for _ in range(2):
        hilo = plan.run(c["method"], c["nsamples"], 0, i0, i0 + cnt)
    st = plan.stats()
    print(kern, cfg, cnt, "i0=%d" % i0, "slow/sample %.2f" % (st["slow_path"] / cnt), hilo, "maxmin %.3f ms dirgen %.3f ms  %.1f Gsp/s" % (st["maxmin_ms"], st["dirgen_ms"], cnt * plan.npoints / st["maxmin_ms"] / 1e6))
