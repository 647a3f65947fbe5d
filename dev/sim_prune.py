"""CPU simulation of the block-pruned search (dev tool; uses the oracle only to make directions).
For a sample of lattice directions: survivors per direction of different block orders / sizes under
(a) thresholds from the warm-up best, (b) thresholds from the final best."""
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
S = int(os.environ.get("S", 512))


def dirs(i0, cnt):
    r = np.zeros((cnt, d))
    L.orc_hw_directions(d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(cnt), P(r))
    return r


rng = np.random.default_rng(0)
# probe winners (random directions on the sphere)
z = np.abs(rng.normal(size=(4096, d))); z /= np.linalg.norm(z, axis=1, keepdims=True)
pr = 1.0 / z
wins = np.zeros(n, int)
for i in range(0, 4096, 256):
    v = (pts[None, :, :] * pr[i:i + 256, None, :]).min(axis=2)
    np.add.at(wins, v.argmax(axis=1), 1)
order_w = np.argsort(-wins, kind="stable")
NPIV = int(os.environ.get("NPIV", 64))
piv = order_w[:NPIV]

colmax = pts.max(axis=0)
norm = pts / colmax
wk = norm.argmin(axis=1); wv = norm.min(axis=1)


def order_weakest():
    return np.lexsort((wv, wk))


def order_weakest2():
    srt = np.argsort(norm, axis=1)
    k1, k2 = srt[:, 0], srt[:, 1]
    return np.lexsort((wv, k2, k1))


def order_kd(leaf):
    idx = np.arange(n)
    out = []

    def rec(ix):
        if len(ix) <= leaf:
            out.append(ix); return
        sub = norm[ix]
        k = (sub.max(axis=0) - sub.min(axis=0)).argmax()
        o = ix[np.argsort(sub[:, k], kind="stable")]
        # split at a multiple of leaf
        h = (len(o) // 2 + leaf - 1) // leaf * leaf
        if h >= len(o): h = len(o) // 2
        rec(o[:h]); rec(o[h:])
    rec(idx)
    return np.concatenate(out)


def order_angular(leaf):
    # kd split on unit direction of the point (p/|p|)
    u = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    idx = np.arange(n); out = []

    def rec(ix):
        if len(ix) <= leaf:
            out.append(ix); return
        sub = u[ix]
        k = (sub.max(axis=0) - sub.min(axis=0)).argmax()
        o = ix[np.argsort(sub[:, k], kind="stable")]
        h = (len(o) // 2 + leaf - 1) // leaf * leaf
        if h >= len(o): h = len(o) // 2
        rec(o[:h]); rec(o[h:])
    rec(idx)
    return np.concatenate(out)


def blocks(order, bs):
    nb = (n + bs - 1) // bs
    padded = np.zeros((nb * bs, d))
    padded[:n] = pts[order]
    return padded.reshape(nb, bs, d).max(axis=1)


windows = [0, N // 4, N // 2, (9 * N) // 10]
R = {w: dirs(w, S) for w in windows}
orders = {"weakest": order_weakest(), "weakest2": order_weakest2(), "kd32": order_kd(32), "ang32": order_angular(32), "kd8": order_kd(8)}
for w in windows:
    r = R[w]
    v = (pts[None, :, :] * r[:, None, :]).min(axis=2)          # S x n
    final = v.max(axis=1)
    b0 = v[:, piv].max(axis=1)
    hit = (b0 == final).mean()
    pcand0 = (v > b0[:, None]).sum(axis=1).mean()
    print(f"window {w}: warm-up hits winner {hit:.3f}; points beating warm-up best: mean {pcand0:.2f}")
    for name, o in orders.items():
        for bs in (32, 16, 8):
            B = blocks(o, bs)
            bb = (B[None, :, :] * r[:, None, :]).min(axis=2)    # bound value per (dir, block)
            s0 = (bb > b0[:, None]).sum(axis=1)
            s1 = (bb > final[:, None]).sum(axis=1)
            print(f"   {name:9s} bs={bs:2d} nb={B.shape[0]:5d}: survivors/dir warm {s0.mean():7.2f} (pts {s0.mean()*bs:7.1f})  final {s1.mean():7.2f} (pts {s1.mean()*bs:7.1f})")
