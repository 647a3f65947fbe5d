"""Synthetic inputs of the BASELINE.json configs, shared by tests, bench.py and
tests/golden/make_golden.py.  Generators follow moocore.generate_ndset's formulas
(python/src/moocore/_moocore.py:1322-1330): |N(0,1)| normalised for sphere fronts,
exponential normalised for simplex fronts, uniform for the random set."""
from __future__ import annotations

import hashlib

import numpy as np


def front(kind: str, n: int, d: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "sphere":
        x = np.abs(rng.normal(size=(n, d)))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "simplex":
        x = rng.exponential(size=(n, d))
        x /= x.sum(axis=1, keepdims=True)
    elif kind == "random":
        x = rng.random((n, d))
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x)


# name -> (method, kind, n, d, nsamples, seed_of_points, mc_seed)
CONFIGS = {
    "cfg1": dict(method="DZ2019-HW", kind="sphere", n=100, d=3, nsamples=100_000, pseed=1, seed=0),
    "cfg2": dict(method="DZ2019-MC", kind="simplex", n=1000, d=5, nsamples=1_000_000, pseed=2, seed=42),
    "cfg3": dict(method="DZ2019-HW", kind="sphere", n=10_000, d=10, nsamples=100_000_000, pseed=3, seed=0),
    "cfg4": dict(method="DZ2019-HW", kind="random", n=50_000, d=20, nsamples=1_000_000_000, pseed=4, seed=0),
    "cfg5": dict(method="DZ2019-MC", kind="random", n=100_000, d=50, nsamples=1_000_000_000, pseed=5, seed=42),
}
REF_VALUE = 1.1


def config_inputs(name: str):
    """(points, ref, maximise) of a config.  cfg5 mixes objectives: every third column is
    maximised (x -> 1 - x, ref -0.1), the others minimised with ref 1.1."""
    c = CONFIGS[name]
    x = front(c["kind"], c["n"], c["d"], c["pseed"])
    d = c["d"]
    ref = np.full(d, REF_VALUE)
    maxi = np.zeros(d, dtype=bool)
    if name == "cfg5":
        cols = np.arange(0, d, 3)
        x[:, cols] = 1.0 - x[:, cols]
        maxi[cols] = True
        ref[cols] = -0.1
    return x, ref, maxi


def transformed(points: np.ndarray, ref: np.ndarray, maxi: np.ndarray) -> np.ndarray:
    """numpy restatement of transform_and_filter (c/hvapprox.c:30-66)."""
    t = np.where(maxi[None, :], points - ref[None, :], ref[None, :] - points)
    keep = (t > 0).all(axis=1)
    return np.ascontiguousarray(t[keep])


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
