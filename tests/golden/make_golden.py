#!/usr/bin/env python3
"""tests/golden/make_golden.py — produce tests/golden/golden.json from the COMPILED REFERENCE.

Run in the build container only (needs oracle/_ref/*.so, i.e. /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py [--full-cfg3]

What is recorded (all numbers as float.hex() so that they round-trip bit-exactly):
  kat        the reference's own known-answer inputs (python/src/moocore/_moocore.py:1090-1100,
             r/tests/testthat/test-doctest-hv_approx.R:9-11) re-evaluated with the reference
  cfg1/cfg2  full estimates of all three methods + exact fpli_hv cross-check values
  cfg3/cfg4  DZ2019-HW window sums and the first per-sample values of every window: start, middle,
             end of the index range and windows around zero-residue lattice indices
  cfg3_full  (only with --full-cfg3; ~35 core-minutes) the reference's hv_approx_hua_wang at the
             full N = 1e8, plus 100 window sums of 1e6 samples each
  dirs       reference directions for a few indices per (d, N)
  digests    sha256 prefixes of the generated inputs, so a numpy stream change is detected
"""
import argparse
import ctypes
import json
import math
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


def load():
    R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhvref.so"))
    for nm in ("hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus",
               "ref_hw_window", "ref_mc_window", "ref_rphi_window"):
        getattr(R, nm).restype = ctypes.c_double
    E = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhvexact.so"))
    E.fpli_hv.restype = ctypes.c_double
    return R, E


def ref_estimate(R, method, pts, ref, maxi, N, seed=0):
    pts = np.ascontiguousarray(pts, dtype=float)
    n, d = pts.shape
    ref = np.ascontiguousarray(np.broadcast_to(np.asarray(ref, float), (d,)))
    mx = np.ascontiguousarray(np.broadcast_to(np.asarray(maxi, bool), (d,))).view(np.uint8)
    a = (P(pts), ctypes.c_size_t(n), ctypes.c_uint8(d), P(ref), mx.ctypes.data_as(u8p), ctypes.c_uint64(N))
    if method == "DZ2019-MC":
        return R.hv_approx_normal(*a, ctypes.c_uint32(seed))
    if method == "DZ2019-HW":
        return R.hv_approx_hua_wang(*a)
    return R.hv_approx_rphi_fang_wang_plus(*a)


def hw_window(args):
    name, i0, cnt, keep = args
    R, _ = load()
    c = hvcases.CONFIGS[name]
    x, ref, maxi = hvcases.config_inputs(name)
    tp = hvcases.transformed(x, ref, maxi)
    vals = np.zeros(cnt)
    s = R.ref_hw_window(P(tp), ctypes.c_size_t(tp.shape[0]), c["d"], ctypes.c_uint64(c["nsamples"]),
                        ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(vals))
    return dict(i0=int(i0), count=int(cnt), sum=float(s).hex(), values=[float(v).hex() for v in vals[:keep]])


def zero_residue_indices(a, N, limit=6):
    """indices i (0-based) with (i+1)*a_k = 0 mod N for some k>=1: i+1 = t*N/gcd."""
    out = []
    for k in range(1, len(a)):
        g = math.gcd(int(a[k]), N)
        if g > 1:
            step = N // g
            for t in range(1, min(g, 3)):
                out.append(t * step - 1)
    return sorted(set(out))[:limit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-cfg3", action="store_true")
    ap.add_argument("--out", default=os.path.join(HERE, "golden.json"))
    args = ap.parse_args()
    R, E = load()
    G = {}
    if os.path.exists(args.out):
        G = json.load(open(args.out))

    # ---- KATs
    x = np.array([[5, 5], [4, 6], [2, 7], [7, 4]], float)
    G["kat2d"] = {m: float(ref_estimate(R, m, x, [10, 10], False, 262144, 42)).hex()
                  for m in ("DZ2019-MC", "DZ2019-HW", "Rphi-FWE+")}
    G["kat2d_doctest"] = {"DZ2019-MC": 38.000806, "DZ2019-HW": 37.999958, "Rphi-FWE+": 37.999979}

    # ---- digests
    G["digests"] = {}
    for name in hvcases.CONFIGS:
        if name == "cfg5":
            continue
        x, ref, maxi = hvcases.config_inputs(name)
        G["digests"][name] = hvcases.digest(x)

    # ---- cfg1 / cfg2 full + exact
    for name in ("cfg1", "cfg2"):
        c = hvcases.CONFIGS[name]
        x, ref, maxi = hvcases.config_inputs(name)
        ent = {}
        for m in ("DZ2019-MC", "DZ2019-HW", "Rphi-FWE+"):
            ent[m] = float(ref_estimate(R, m, x, ref, maxi, c["nsamples"], 42)).hex()
        ent["exact"] = float(E.fpli_hv(P(x), ctypes.c_size_t(x.shape[0]), ctypes.c_uint8(c["d"]), P(ref))).hex()
        if name == "cfg2":
            x4 = np.ascontiguousarray(x[:, :4])
            ent["exact_4d_slice"] = float(E.fpli_hv(P(x4), ctypes.c_size_t(x4.shape[0]), ctypes.c_uint8(4), P(ref[:4].copy()))).hex()
            for m in ("DZ2019-MC", "DZ2019-HW"):
                ent["slice4_" + m] = float(ref_estimate(R, m, x4, ref[:4], maxi[:4], c["nsamples"], 42)).hex()
        G[name] = ent
        print(name, {k: float.fromhex(v) for k, v in ent.items()})

    # ---- cfg3 / cfg4 windows
    jobs = []
    for name, cnt in (("cfg3", 1500), ("cfg4", 300)):
        c = hvcases.CONFIGS[name]
        N = c["nsamples"]
        a = (ctypes.c_uint64 * 64)()
        R.ref_hw_polar_a(c["d"], ctypes.c_uint64(N), a)
        zr = zero_residue_indices([a[k] for k in range(c["d"] - 1)], N)
        starts = [0, N // 7, N // 2, N - cnt] + [max(0, z - cnt // 2) for z in zr]
        G[name + "_zero_residue"] = [int(z) for z in zr]
        for s in starts:
            jobs.append((name, int(s), cnt, 64))
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(hw_window, jobs))
    print("windows done in %.1fs" % (time.time() - t0))
    for (name, *_), r in zip(jobs, res):
        G.setdefault(name + "_windows", [])
    for name in ("cfg3", "cfg4"):
        G[name + "_windows"] = [r for (n_, *_), r in zip(jobs, res) if n_ == name]

    # ---- reference directions
    G["dirs"] = []
    for d, N in ((3, 100_000), (5, 1_000_000), (10, 100_000_000), (20, 1_000_000_000)):
        for i0 in (0, N // 3, N - 8):
            r = np.zeros((8, d))
            R.ref_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(8), P(r))
            G["dirs"].append(dict(d=d, N=N, i0=int(i0), r=[[float(v).hex() for v in row] for row in r]))

    # ---- MC / Rphi stream heads
    z = np.zeros(4096)
    R.ref_mc_normals(ctypes.c_uint32(42), ctypes.c_uint64(4096), P(z))
    G["mc_normals_seed42"] = [float(v).hex() for v in z[:64]]
    G["mc_normals_seed42_sum4096"] = float(z.sum()).hex()

    if args.full_cfg3:
        name = "cfg3"
        c = hvcases.CONFIGS[name]
        W = 1_000_000
        jobs = [(name, i * W, W, 0) for i in range(c["nsamples"] // W)]
        t0 = time.time()
        with ProcessPoolExecutor(max_workers=8) as ex:
            res = list(ex.map(hw_window, jobs))
        G["cfg3_full_windows"] = [r["sum"] for r in res]
        print("cfg3 full windows in %.1fs" % (time.time() - t0))
        json.dump(G, open(args.out, "w"), indent=0)
        t0 = time.time()
        x, ref, maxi = hvcases.config_inputs(name)
        G["cfg3_full"] = float(ref_estimate(R, "DZ2019-HW", x, ref, maxi, c["nsamples"])).hex()
        print("cfg3 full single run in %.1fs: %r" % (time.time() - t0, float.fromhex(G["cfg3_full"])))

    json.dump(G, open(args.out, "w"), indent=0)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
