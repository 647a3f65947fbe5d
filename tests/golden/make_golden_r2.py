#!/usr/bin/env python3
"""tests/golden/make_golden_r2.py — produce tests/golden/golden_r2.json from the COMPILED REFERENCE.

Run in the build container only (needs oracle/_ref/*.so, i.e. /root/reference):
    make -C oracle ref && python tests/golden/make_golden_r2.py

What is recorded (float.hex() everywhere):
  cfg5_windows   BASELINE config 5 (DZ2019-MC seed 42, d = 50, n = 1e5, every third objective maximised) lies
                 outside the reference's domain (d > 31): these come from the PATCHED reference
                 oracle/_ref/libhvref_d64.so (oracle/patch_ref_d64.py: dimension limit 64, c_d and 1/phi_d tables
                 extended by their generator formulas; everything else is the reference).  Window sums, the first
                 per-sample values and the first reciprocal directions at three depths of the MT19937 stream
                 (sample 0, ~1e6, ~1e7: 5e7 / 5e8 normals in).
  cfg5_small     full hv_approx_normal / hv_approx_rphi_fang_wang_plus of the patched reference on the first 2000
                 points of config 5 with 20 000 samples (pins c_50 and the Rphi tables past d = 31)
  hw_d21_d22     DZ2019-HW at d = 21 and 22, the last dimensions where the reference terminates:
                 directions, per-sample values and window sums from the unmodified reference
  ext_dims       MC / Rphi window sums at d = 33, 40, 64 from the patched reference
"""
import ctypes
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hvcases  # noqa: E402

dp = ctypes.POINTER(ctypes.c_double)
u8p = ctypes.POINTER(ctypes.c_uint8)


def P(a):
    return a.ctypes.data_as(dp)


def load(name):
    R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", name))
    for nm in ("hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus",
               "ref_hw_window", "ref_mc_window", "ref_rphi_window"):
        getattr(R, nm).restype = ctypes.c_double
    return R


def hexl(a):
    return [float(v).hex() for v in a]


def cfg5_window(args):
    i0, cnt, keep = args
    R = load("libhvref_d64.so")
    c = hvcases.CONFIGS["cfg5"]
    x, ref, maxi = hvcases.config_inputs("cfg5")
    tp = hvcases.transformed(x, ref, maxi)
    vals = np.zeros(cnt)
    r = np.zeros((cnt, c["d"]))
    s = R.ref_mc_window(P(tp), ctypes.c_size_t(tp.shape[0]), c["d"], ctypes.c_uint32(c["seed"]),
                        ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(vals), P(r))
    return dict(i0=int(i0), count=int(cnt), npoints=int(tp.shape[0]), sum=float(s).hex(), values=hexl(vals[:keep]),
                dirs=[hexl(row) for row in r[:4]])


def main():
    out = os.path.join(HERE, "golden_r2.json")
    G = {}
    R64 = load("libhvref_d64.so")
    R = load("libhvref.so")

    x, ref, maxi = hvcases.config_inputs("cfg5")
    G["cfg5_digest"] = hvcases.digest(x)
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=3) as ex:
        G["cfg5_windows"] = list(ex.map(cfg5_window, [(0, 256, 64), (1_000_003, 256, 64), (10_000_019, 256, 64)]))
    print("cfg5 windows in %.1fs" % (time.time() - t0))

    # full entry points of the patched reference at d = 50 (c_d and Rphi tables past 31)
    xs = np.ascontiguousarray(x[:2000])
    mx = maxi.view(np.uint8)
    a = (P(xs), ctypes.c_size_t(xs.shape[0]), ctypes.c_uint8(50), P(ref), mx.ctypes.data_as(u8p), ctypes.c_uint64(20000))
    G["cfg5_small"] = {"n": 2000, "nsamples": 20000,
                       "DZ2019-MC": float(R64.hv_approx_normal(*a, ctypes.c_uint32(42))).hex(),
                       "Rphi-FWE+": float(R64.hv_approx_rphi_fang_wang_plus(*a)).hex()}
    print("cfg5_small", {k: (float.fromhex(v) if isinstance(v, str) else v) for k, v in G["cfg5_small"].items()})

    # MC / Rphi windows at other extension dimensions
    G["ext_dims"] = []
    for d in (33, 40, 64):
        tp = np.ascontiguousarray(1.1 - hvcases.front("random", 131, d, 100 + d))
        v1, v2 = np.zeros(300), np.zeros(300)
        s1 = R64.ref_mc_window(P(tp), ctypes.c_size_t(131), d, ctypes.c_uint32(7), ctypes.c_uint64(1000), ctypes.c_uint64(300), P(v1), None)
        s2 = R64.ref_rphi_window(P(tp), ctypes.c_size_t(131), d, ctypes.c_uint64(1000), ctypes.c_uint64(300), P(v2), None, None)
        G["ext_dims"].append(dict(d=d, n=131, pseed=100 + d, i0=1000, count=300, mc_seed=7, mc_sum=float(s1).hex(),
                                  mc_values=hexl(v1[:32]), rphi_sum=float(s2).hex(), rphi_values=hexl(v2[:32])))

    # DZ2019-HW at the edge of the reference's working range
    G["hw_d21_d22"] = []
    for d, N in ((21, 1_000_000), (22, 1_000_000), (21, 100_000_000), (22, 1_000_000_000)):
        tp = np.ascontiguousarray(1.1 - hvcases.front("sphere", 200, d, d))
        for i0 in (0, N // 3, N - 500):
            vals = np.zeros(500)
            r = np.zeros((8, d))
            s = R.ref_hw_window(P(tp), ctypes.c_size_t(200), d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(500), P(vals))
            R.ref_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(8), P(r))
            G["hw_d21_d22"].append(dict(d=d, N=N, n=200, pseed=d, i0=int(i0), count=500, sum=float(s).hex(),
                                        values=hexl(vals[:64]), dirs=[hexl(row) for row in r]))
    json.dump(G, open(out, "w"), indent=0)
    print("wrote", out)


if __name__ == "__main__":
    main()
