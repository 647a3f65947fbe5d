#!/usr/bin/env python3
"""tests/golden/make_golden_datasets.py — the dataset known-answers of the reference's doctest
(python/src/moocore/_moocore.py:1101-1130) as input/output vectors.

Run in the build container only (needs /root/reference and oracle/_ref/libhvref.so):
    make -C oracle ref && python tests/golden/make_golden_datasets.py

The two packaged datasets are read from where they lie, preprocessed exactly like the doctest does
(merge the sets; filter_dominated; normalise to [1, 2]) with plain numpy, evaluated by the COMPILED REFERENCE, checked
against the numbers printed in the doctest, and written — processed inputs and reference outputs, float.hex() —
to tests/golden/golden_datasets.json."""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import load, ref_estimate  # noqa: E402

DATA = "/root/reference/python/src/moocore/data"


def read_sets(path):
    rows = [[float(v) for v in ln.split()] for ln in open(path) if ln.strip() and not ln.startswith("#")]
    return np.array(rows)


def filter_dominated(x, maximise=False):
    """Rows not dominated by another row (weak dominance with one strict inequality); duplicates keep the first."""
    y = -x if maximise else x
    keep = np.ones(len(y), bool)
    for i in range(len(y)):
        le = (y <= y[i]).all(axis=1)
        lt = (y < y[i]).any(axis=1)
        dom = le & lt
        dup = (y == y[i]).all(axis=1)
        dup[i:] = False
        keep[i] = not (dom.any() or dup.any())
    return x[keep]


def main():
    R, E = load()
    N = 262144
    out = {}
    x = read_sets(os.path.join(DATA, "input1.dat"))
    for name, pts in (("input1", x), ("input1_filtered", filter_dominated(x))):
        rec = {"points": [[v.hex() for v in row] for row in pts.tolist()], "ref": 10.0, "maximise": False, "nsamples": N}
        for m in ("Rphi-FWE+", "DZ2019-HW"):
            rec[m] = float(ref_estimate(R, m, pts, 10.0, False, N)).hex()
            assert abs(float.fromhex(rec[m]) - 93.5533) < 5e-5, (name, m, float.fromhex(rec[m]))
        out[name] = rec
    x = read_sets(os.path.join(DATA, "ran.10pts.9d.10"))
    x = filter_dominated(-x, maximise=True)
    lo, hi = x.min(axis=0), x.max(axis=0)
    x = 1.0 + (x - lo) / (hi - lo)
    rec = {"points": [[v.hex() for v in row] for row in x.tolist()], "ref": 0.9, "maximise": True, "nsamples": N}
    want = {"Rphi-FWE+": 0.48397532301, "DZ2019-HW": 0.483083547763}
    for m in want:
        rec[m] = float(ref_estimate(R, m, x, 0.9, True, N)).hex()
        assert abs(float.fromhex(rec[m]) - want[m]) < 1e-11, (m, float.fromhex(rec[m]), want[m])
    out["ran10pts9d_normalised"] = rec
    with open(os.path.join(HERE, "golden_datasets.json"), "w") as f:
        json.dump(out, f, indent=0)
    for k, v in out.items():
        print(k, len(v["points"]), "points:", {m: float.fromhex(v[m]) for m in ("Rphi-FWE+", "DZ2019-HW")})


if __name__ == "__main__":
    main()
