"""Cost of describing one rank's pair range of the DZ2019-MC stream (cfg5 on 8 GPUs: 6.4e9 pairs per rank) against the
prefix parse it replaces (rank g of G used to parse g/G of the whole stream)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
from moocore_b200._api import mc_segment_layout
N, d, world = 10**9, 50, 8
x = hvcases.front("random", 64, d, 1)
with mb.Plan(x, np.full(d, 1.1), False, allow_extension=True) as plan:
    for g in (0, 1, 6):
        p0, npairs = mc_segment_layout(N, d, world, g)
        for frac in (0.1, 1.0):
            t0 = time.time()
            seg = plan.mc_segment_scan(42, p0, int(npairs * frac))
            dt = time.time() - t0
            print("rank %d  pair0 %.3e  npairs %.3e  scan %.3f s  (%.2f Gpairs/s)  merge %d  skip_out %d  body %d" % (
                g, p0, npairs * frac, dt, npairs * frac / dt / 1e9, seg.merge, seg.skip_out, seg.body), flush=True)
    # the prefix parse of the same number of normals, for comparison
    t0 = time.time()
    plan.run("DZ2019-MC", N, 42, 2_000_000, 2_000_001)
    print("prefix parse of 1e8 normals: %.3f s" % (time.time() - t0))
