"""Seeded inputs of the whv_hype parity cases (shared by tests/golden/make_golden_whv.py and tests/test_whv_hype.py)."""
import numpy as np

DISTS = ("uniform", "exponential", "point")

# name -> (npoints, nsamples, seed of the inputs, MT19937 seed)
CASES = {
    "one_point": (1, 100000, 1, 42),
    "small": (7, 5000, 2, 12345),
    "front100": (100, 100000, 3, 7),
    "cloud1000": (1000, 50000, 4, 2**32 - 2),
    "odd_sizes": (333, 33333, 5, 0),
    "duplicates": (64, 20001, 6, 99),
}


def case_inputs(name):
    """points (n x 2, minimisation), ideal, ref, mu of the exponential, goal point of the normal distribution."""
    n, nsamples, iseed, seed = CASES[name]
    rng = np.random.default_rng(iseed)
    if name == "front100":
        t = np.sort(rng.random(n))
        pts = np.stack([1.0 + 2.5 * t, 1.0 + 2.5 * (1.0 - t) ** 2], axis=1)
    elif name == "duplicates":
        base = 1.0 + 3.0 * rng.random((8, 2))
        pts = base[rng.integers(0, 8, n)]
    else:
        pts = 0.5 + 3.5 * rng.random((n, 2))          # some points outside [ideal, ref]
    ideal = np.array([1.0, 1.0]) - 0.25 * rng.random(2)
    ref = np.array([4.0, 4.0]) + 0.25 * rng.random(2)
    mu_expo = float(0.05 + 0.4 * rng.random())
    goal = 1.5 + 2.0 * rng.random(2)
    return np.ascontiguousarray(pts), ideal, ref, mu_expo, goal, nsamples, seed
