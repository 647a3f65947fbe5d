#!/usr/bin/env python3
"""tests/golden/make_golden_contrib.py — exact hypervolume contributions from the COMPILED REFERENCE
(hv_contributions, c/hv_contrib.c:348, in oracle/_ref/libhvexact.so) for the cross-check of the approximate
contributions of moocore_b200.hv_contributions_approx.  Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_contrib.py"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hvcases  # noqa: E402

dp = ctypes.POINTER(ctypes.c_double)


def cases():
    """name -> (points, ref): fronts with a few dominated rows, a duplicate and a row that does not dominate ref"""
    out = {}
    for name, kind, n, d, seed in (("sphere3", "sphere", 60, 3, 31), ("simplex4", "simplex", 40, 4, 32), ("sphere2", "sphere", 25, 2, 33)):
        x = hvcases.front(kind, n, d, seed)
        extra = np.vstack([x[:5] + 0.03, x[7:8], np.full((1, d), 1.3)])      # dominated, duplicate of row 7, outside ref
        out[name] = (np.ascontiguousarray(np.vstack([x, extra])), np.full(d, 1.1))
    return out


def main():
    E = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhvexact.so"))
    E.hv_contributions.restype = ctypes.c_double
    E.fpli_hv.restype = ctypes.c_double
    G = {}
    for name, (x, ref) in cases().items():
        n, d = x.shape
        keep = (x < ref).all(axis=1)                 # hv_contributions wants points that dominate ref
        xs = np.ascontiguousarray(x[keep])
        ent = dict(digest=hvcases.digest(x))
        for key, ign in (("contributions", True), ("contributions_keep_dominated", False)):
            hvc = np.zeros(xs.shape[0])
            pts = xs.copy()
            E.hv_contributions(hvc.ctypes.data_as(dp), pts.ctypes.data_as(dp), ctypes.c_size_t(xs.shape[0]), ctypes.c_uint8(d),
                               ref.ctypes.data_as(dp), ctypes.c_bool(ign))
            full = np.zeros(n)
            full[np.nonzero(keep)[0]] = hvc
            ent[key] = [float(v).hex() for v in full]
            print(name, key, "sum", full.sum(), "zeros", int((full == 0).sum()))
        hv = E.fpli_hv(xs.ctypes.data_as(dp), ctypes.c_size_t(xs.shape[0]), ctypes.c_uint8(d), ref.ctypes.data_as(dp))
        ent["hv"] = float(hv).hex()
        G[name] = ent
    json.dump(G, open(os.path.join(HERE, "golden_contrib.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
