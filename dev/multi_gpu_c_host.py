#!/usr/bin/env python3
"""dev/multi_gpu_c_host.py — the C-host multi-GPU path on hardware.

One process, no torch: binds `hv_approx_hua_wang` (the symbol of c/hvapprox.h:25-28) from libmoocore_b200.so through
ctypes exactly as a Python/R caller of the reference would, sets HVB200_NGPUS=k, and calls it.  Inside, the library
runs one host thread per GPU, block-cyclic direction shards, and one ncclAllGather of the (hi, lo) pairs
(hvb200_hv_approx_multi -> hvb200cu_nccl_allgather; libnccl through dlopen).

Prints, per k: the estimate, its relative difference to the 1-GPU estimate, whether the exchange went through NCCL,
wall time per call (host buffers in, double out) and the efficiency against k x the 1-GPU call.  Run with
NCCL_DEBUG=INFO to get the communicator lines (kept under profiles/).

    python dev/multi_gpu_c_host.py [--config cfg3] [--reps 5] [--ngpus 1,2,4,8] [--json out.json]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hvcases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ngpus", default="")
    ap.add_argument("--nsamples", type=int, default=0)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    L = ctypes.CDLL(os.path.join(ROOT, "moocore_b200", "libmoocore_b200.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    for nm in ("hv_approx_hua_wang", "hv_approx_normal", "hv_approx_rphi_fang_wang_plus"):
        getattr(L, nm).restype = ctypes.c_double
    L.hvb200_last_error.restype = ctypes.c_char_p
    navail = L.hvb200_device_count()
    ks = [int(v) for v in args.ngpus.split(",")] if args.ngpus else [k for k in (1, 2, 4, 8) if k <= navail]
    c = hvcases.CONFIGS[args.config]
    N = args.nsamples or c["nsamples"]
    x, ref, maxi = hvcases.config_inputs(args.config)
    mx = maxi.view(np.uint8)
    a = (x.ctypes.data_as(dp), ctypes.c_size_t(x.shape[0]), ctypes.c_uint8(c["d"]), ref.ctypes.data_as(dp),
         mx.ctypes.data_as(u8p), ctypes.c_uint64(N))

    def call():
        if c["method"] == "DZ2019-HW":
            return L.hv_approx_hua_wang(*a)
        if c["method"] == "DZ2019-MC":
            return L.hv_approx_normal(*a, ctypes.c_uint32(c["seed"]))
        return L.hv_approx_rphi_fang_wang_plus(*a)

    npts = hvcases.transformed(x, ref, maxi).shape[0]
    rows = []
    base = None
    for k in ks:
        os.environ["HVB200_NGPUS"] = str(k)
        t0 = time.perf_counter()
        est = call()                      # first call: CUDA contexts, NCCL communicator, workspace allocation
        first = time.perf_counter() - t0
        if est != est:
            raise SystemExit(f"k={k}: call failed: {L.hvb200_last_error().decode()}")
        times = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            e2 = call()
            times.append(time.perf_counter() - t0)
            assert e2 == est, "the estimate must not depend on the run"
        exch = L.hvb200_multi_last_exchange()
        t = float(np.median(times))
        if base is None:
            base = (k, est, t)
        row = dict(ngpus=k, estimate=est, rel_diff_vs_first=abs(est - base[1]) / abs(base[1]),
                   exchange={0: "none", 1: "nccl", 2: "host"}[exch], first_call_s=first, ms_per_call=1e3 * t,
                   ms_min=1e3 * min(times), gsample_point_per_s=N * npts / t / 1e9,
                   efficiency_vs_first=(base[2] * base[0]) / (t * k))
        rows.append(row)
        print(json.dumps(row), flush=True)
        assert row["rel_diff_vs_first"] <= 1e-13, row
        assert k == 1 or exch == 1 or os.environ.get("HVB200_NO_NCCL"), "NCCL exchange expected"
    if args.json:
        json.dump(dict(config=args.config, nsamples=N, npoints=npts, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
