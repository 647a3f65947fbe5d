"""CPU simulation (dev tool): quality of the phase-A lower bound b0 and the surviving blocks for
  (a) the best of the NPIV global pivots (what k_maxmin_bp does), against
  (b) a two-level choice: best of C coarse pivots, then the best of that pivot's M-entry neighbour list (the points that win
      the probe directions in the coarse pivot's cell)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hvcases
from util import P
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "libhvoracle.so"))
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
c = hvcases.CONFIGS[cfg]
x, ref, maxi = hvcases.config_inputs(cfg)
pts = hvcases.transformed(x, ref, maxi)
n, d = pts.shape
N = c["nsamples"]
S = int(os.environ.get("S", 2048))
NPROBE = int(os.environ.get("NPROBE", 65536))

def hw_dirs(i0, cnt):
    r = np.zeros((cnt, d))
    L.orc_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(r))
    return r

def maxmin(P_, r):            # values [len(r), len(P_)] in chunks
    out_arg = np.zeros(len(r), int); out_val = np.zeros(len(r))
    for i in range(0, len(r), 128):
        v = (P_[None, :, :] * r[i:i + 128, None, :]).min(axis=2)
        out_arg[i:i + 128] = v.argmax(axis=1); out_val[i:i + 128] = v.max(axis=1)
    return out_arg, out_val

rng = np.random.default_rng(0)
z = np.abs(rng.normal(size=(NPROBE, d))); z /= np.linalg.norm(z, axis=1, keepdims=True)
pr = 1.0 / z
parg, pval = maxmin(pts, pr)
wins = np.bincount(parg, minlength=n)
order_w = np.argsort(-wins, kind="stable")
print("distinct probe winners:", (wins > 0).sum(), "of", n)

# kd order, 32-point blocks, exact bounds
colmax = pts.max(axis=0); norm = pts / colmax
def order_kd(leaf):
    out = []
    def rec(ix):
        if len(ix) <= leaf: out.append(ix); return
        sub = norm[ix]; k = (sub.max(axis=0) - sub.min(axis=0)).argmax()
        o = ix[np.argsort(sub[:, k], kind="stable")]
        h = (len(o) // 2 + leaf - 1) // leaf * leaf
        if h >= len(o): h = len(o) // 2
        rec(o[:h]); rec(o[h:])
    rec(np.arange(n)); return np.concatenate(out)
ordk = order_kd(32)
nb = (n + 31) // 32
bmax = np.zeros((nb, d))
for b in range(nb): bmax[b] = pts[ordk[b * 32:(b + 1) * 32]].max(axis=0)

test = hw_dirs(N // 2, S)
targ, tval = maxmin(pts, test)
def survivors(b0):
    cnt = np.zeros(len(test))
    for i in range(0, len(test), 256):
        bv = (bmax[None, :, :] * test[i:i + 256, None, :]).min(axis=2)
        cnt[i:i + 256] = (bv > b0[i:i + 256, None] * (1 - 2 ** -20)).sum(axis=1)
    return cnt
print("final max: survivors/dir %.2f" % survivors(tval).mean())
for NPIV in (128, 320, 512):
    piv = order_w[:NPIV]
    _, b0 = maxmin(pts[piv], test)
    print("global %4d pivots: b0/true mean %.5f  exact-hit %.3f  survivors/dir %.2f" % (NPIV, (b0 / tval).mean(), (b0 == tval).mean(), survivors(b0).mean()))
for C, M in ((32, 32), (64, 32), (64, 64), (128, 32), (128, 64), (256, 32)):
    coarse = order_w[:C]
    carg, _ = maxmin(pts[coarse], pr)                     # cell of every probe direction
    lists = np.zeros((C, M), int)
    for cc in range(C):
        w = np.bincount(parg[carg == cc], minlength=n)
        lists[cc] = np.argsort(-w, kind="stable")[:M]
        if (carg == cc).sum() == 0: lists[cc] = order_w[:M]
    tc, b_c = maxmin(pts[coarse], test)
    b0 = b_c.copy()
    for i in range(len(test)):
        v = (pts[lists[tc[i]]] * test[i][None, :]).min(axis=1).max()
        b0[i] = max(b0[i], v)
    print("two-level C=%3d M=%3d (%3d evals): b0/true mean %.5f  exact-hit %.3f  survivors/dir %.2f" % (C, M, C + M, (b0 / tval).mean(), (b0 == tval).mean(), survivors(b0).mean()))
