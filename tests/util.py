import ctypes

import numpy as np

dp = ctypes.POINTER(ctypes.c_double)
u8p = ctypes.POINTER(ctypes.c_uint8)


def P(a):
    return a.ctypes.data_as(dp)


def fh(s):
    return float.fromhex(s)


def c_args(pts, ref, maxi, N):
    pts = np.ascontiguousarray(pts, dtype=float)
    n, d = pts.shape
    ref = np.ascontiguousarray(np.broadcast_to(np.asarray(ref, float), (d,)))
    mx = np.ascontiguousarray(np.broadcast_to(np.asarray(maxi, bool), (d,))).view(np.uint8)
    keep = (pts, ref, mx)
    return keep, (P(pts), ctypes.c_size_t(n), ctypes.c_uint8(d), P(ref), mx.ctypes.data_as(u8p), ctypes.c_uint64(N))


def estimate(L, prefix, method, pts, ref, maxi, N, seed=0):
    keep, a = c_args(pts, ref, maxi, N)
    if method == "DZ2019-MC":
        return getattr(L, prefix + "hv_approx_normal")(*a, ctypes.c_uint32(seed))
    if method == "DZ2019-HW":
        return getattr(L, prefix + "hv_approx_hua_wang")(*a)
    return getattr(L, prefix + "hv_approx_rphi_fang_wang_plus")(*a)


def rel(a, b):
    return abs(a - b) / abs(b)
