"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hvcases, moocore_b200 as mb
x = hvcases.front("sphere", 3000, 6, 1)
for kern in ("exact", "screen", "dpx", "pruned"):
    with mb.Plan(x, 1.1, kernel=kern) as plan:
        print(kern, plan.run("DZ2019-HW", 200_000, 0, 90_000, 100_000))
x = hvcases.front("random", 70_000, 20, 2)           # multi-segment pruned kernel, device transform
with mb.Plan(x, 1.1, kernel="pruned") as plan:
    print("pruned seg", plan.run("DZ2019-HW", 1_000_000, 0, 500_000, 503_000))
x = hvcases.front("random", 900, 40, 3)
with mb.Plan(x, 1.1, kernel="pruned", allow_extension=True) as plan:
    print("pruned d=40", plan.run("DZ2019-MC", 3000, 5, 0, 3000))
print("mc", mb.hv_approx(hvcases.front("simplex", 300, 5, 4), ref=1.1, nsamples=400_000, seed=3, method="DZ2019-MC"))
print("rphi", mb.hv_approx(hvcases.front("simplex", 300, 5, 4), ref=1.1, nsamples=50_000, method="Rphi-FWE+"))
print("sets", mb.hv_approx_sets([hvcases.front("sphere", 50 + i, 4, i) for i in range(6)], 1.1, nsamples=5000, method="DZ2019-HW"))
