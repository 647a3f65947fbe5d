#!/usr/bin/env python3
"""tests/golden/make_golden_whv.py — known answers of the reference's HypE estimator (c/whv_hype.c).

Run in the build container only (needs /root/reference and oracle/_ref/libwhvref.so):
    make -C oracle ref && python tests/golden/make_golden_whv.py

Evaluates the COMPILED REFERENCE on the seeded cases of tests/whvcases.py and on the docstring examples of
python/src/moocore/_moocore.py:2832-2864 (checked against the printed numbers), and writes the outputs as float.hex()
to tests/golden/golden_whv.json."""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import whvcases  # noqa: E402

R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libwhvref.so"))
for f in (R.whv_hype_unif, R.whv_hype_expo, R.whv_hype_gaus):
    f.restype = ctypes.c_double
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def ref_whv(dist, pts, ideal, ref, nsamples, seed, mu_expo, goal):
    pts = np.ascontiguousarray(pts, dtype=float)
    ideal = np.ascontiguousarray(ideal, dtype=float)
    ref = np.ascontiguousarray(ref, dtype=float)
    args = (P(pts), pts.shape[0], P(ideal), P(ref), nsamples, ctypes.c_uint32(seed))
    if dist == "uniform":
        return R.whv_hype_unif(*args)
    if dist == "exponential":
        return R.whv_hype_expo(*args, ctypes.c_double(mu_expo))
    g = np.array(goal, dtype=float)           # the reference normalises mu in place
    return R.whv_hype_gaus(*args, P(g))


out = {"cases": {}, "doctest": []}
for name in whvcases.CASES:
    pts, ideal, ref, mu_expo, goal, nsamples, seed = whvcases.case_inputs(name)
    out["cases"][name] = {d: ref_whv(d, pts, ideal, ref, nsamples, seed, mu_expo, goal).hex() for d in whvcases.DISTS}

DOC = [([[2, 2]], "uniform", None, 3.99807), ([[3, 1]], "uniform", None, 3.00555),
       ([[2, 2]], "exponential", 0.2, 1.14624), ([[3, 1]], "exponential", 0.2, 1.66815),
       ([[2, 2]], "point", [2.9, 0.9], 0.64485), ([[3, 1]], "point", [2.9, 0.9], 4.03632)]
for pts, dist, mu, printed in DOC:
    v = ref_whv(dist, np.array(pts, dtype=float), np.full(2, 1.0), np.full(2, 4.0), 100000, 42, mu, mu)
    assert abs(v - printed) < 1e-5, (pts, dist, v, printed)
    out["doctest"].append({"points": pts, "dist": dist, "mu": mu, "printed": printed, "value": v.hex()})

json.dump(out, open(os.path.join(HERE, "golden_whv.json"), "w"), indent=1)
print("wrote golden_whv.json:", {k: float.fromhex(v["uniform"]) for k, v in out["cases"].items()})
