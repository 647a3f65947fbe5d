"""One FULL estimate of a BASELINE config on one GPU, host buffers in, estimate out (plan creation, preparation,
direction generation, max-min, reduction, finalisation all inside the timed call).
    python dev/full_run.py cfg4 [reps] [json-out]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hvcases, moocore_b200 as mb
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = hvcases.CONFIGS[cfg]
x, ref, maxi = hvcases.config_inputs(cfg)
N, d = c["nsamples"], c["d"]
rows = []
for rep in range(reps):
    t0 = time.perf_counter()
    with mb.Plan(x, ref, maxi, allow_extension=d > 31) as plan:
        hilo = plan.run(c["method"], N, c["seed"])
        est = mb.hv_approx_finalize(d, N, hilo)
        st = plan.stats()
        npts = plan.npoints
    dt = time.perf_counter() - t0
    row = dict(config=cfg, rep=rep, estimate=est, estimate_hex=float(est).hex(), seconds=dt, npoints=npts,
               gsample_point_per_s=N * npts / dt / 1e9, maxmin_ms=st["maxmin_ms"], dirgen_ms=st["dirgen_ms"],
               maxmin_launches=st["maxmin_launches"], dirgen_launches=st["dirgen_launches"], slow_per_sample=st["slow_path"] / N,
               newton_capped=st["newton_capped"], kernel=st["kernel"])
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 3:
    json.dump(rows, open(sys.argv[3], "w"), indent=1)
