"""Concurrent calls from several host threads (the reference is re-entrant and CFFI/ctypes release the GIL,
SURVEY §8b "Threading"): every call owns its plan and stream; the shared state — workspace pool, Rphi direction
cache, MT19937 jump tables, transform scratch, constant tables — must be safe under contention."""
import threading

import numpy as np
import pytest

import hvcases

pytestmark = pytest.mark.gpu


def _jobs():
    rng = np.random.default_rng(77)
    jobs = []
    for i in range(48):
        d = int(rng.choice([2, 3, 5, 8, 12]))
        n = int(rng.choice([1, 40, 300, 2500, 40000]))
        method = ["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"][i % 3]
        N = int(rng.choice([1000, 50_000, 300_000, 700_000]))
        x = hvcases.front(["sphere", "simplex", "random"][i % 3], n, d, 1000 + i)
        jobs.append((x, method, N, i))
    # large enough for the block-pruned kernel (plan-time preparation on the host) and the device transform
    for i, (n, d) in enumerate([(3000, 6), (60000, 5), (3000, 6), (5000, 10)]):
        jobs.append((hvcases.front("sphere", n, d, 2000 + i), ["DZ2019-HW", "DZ2019-MC"][i % 2], 1_500_000, i))
    return jobs


def test_concurrent_calls_match_serial_results(mb):
    jobs = _jobs()
    got = [None] * len(jobs)
    errors = []

    def worker(tid):
        try:
            for k in range(tid, len(jobs), 8):
                x, method, N, seed = jobs[k]
                got[k] = mb.hv_approx(x, ref=1.1, nsamples=N, seed=seed, method=method)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k, (x, method, N, seed) in enumerate(jobs):
        want = mb.hv_approx(x, ref=1.1, nsamples=N, seed=seed, method=method)
        assert got[k] == want, (k, method, x.shape, N)


def test_no_device_memory_growth_over_many_calls(mb):
    # plans come and go through the workspace pool; repeated calls of every kind must not keep allocating
    import torch

    def burst():
        rng = np.random.default_rng(3)
        for i in range(60):
            d = int(rng.choice([2, 4, 7]))
            n = int(rng.choice([30, 400, 3000]))
            x = hvcases.front("sphere", n, d, i)
            mb.hv_approx(x, ref=1.1, nsamples=int(rng.choice([2000, 80_000, 700_000])), seed=i,
                         method=["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"][i % 3])
        mb.hv_approx_sets([hvcases.front("sphere", 50, 3, k) for k in range(10)], 1.1, nsamples=5000, method="DZ2019-HW")
        mb.epsilon_additive(rng.random((500, 4)), rng.random((400, 4)))
        mb.igd_plus(rng.random((500, 4)), rng.random((400, 4)))

    burst()                                   # warm: pools, caches and tables reach their steady size
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(3):
        burst()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, (free0, free1)
