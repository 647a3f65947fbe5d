"""Host-side cost breakdown of one end-to-end call (plan create, first run incl. preparation, second run)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hvcases, moocore_b200 as mb
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
c = hvcases.CONFIGS[cfg]
x, ref, maxi = hvcases.config_inputs(cfg)
meth = c["method"] if c["d"] <= 31 else "DZ2019-MC"
for rep in range(3):
    t0 = time.perf_counter()
    plan = mb.Plan(x, ref, maxi, allow_extension=True)
    t1 = time.perf_counter()
    plan.run(meth, c["nsamples"], c["seed"], 0, cnt)
    t2 = time.perf_counter()
    plan.run(meth, c["nsamples"], c["seed"], 0, cnt)
    t3 = time.perf_counter()
    st = plan.stats()
    plan.close()
    t4 = time.perf_counter()
    print("%s rep %d: create %.2f ms, first run %.2f ms, second run %.2f ms (maxmin %.2f dirgen %.2f), close %.2f ms, kernel %d" % (
        cfg, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), st["maxmin_ms"], st["dirgen_ms"], 1e3 * (t4 - t3), st["kernel"]))
