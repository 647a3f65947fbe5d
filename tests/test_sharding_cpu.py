"""Host logic of the multi-GPU path, on CPU: range partition, and a world_size-2 gloo run of
the exchange step (all-gather of (hi, lo) partials + ordered finalisation).  The per-rank
partial sums come from the CPU oracle here (test infrastructure); on the GPU box the same
plumbing carries the kernels' partials (bench.py, tests/test_parity_gpu.py)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from moocore_b200.sharding import gather_partials, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N", [1, 7, 100, 262144, 100_000_000, 2**31])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_ranges_partition(N, world):
    edges = [shard_range(N, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == N
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("N,world,block", [(1, 1, 8), (100, 3, 7), (10_000_019, 8, 1 << 21), (4096, 4, 1024), (5, 8, 2)])
def test_shard_blocks_partition(N, world, block):
    from moocore_b200.sharding import shard_blocks
    ranges = sorted(r for g in range(world) for r in shard_blocks(N, g, world, block))
    assert ranges[0][0] == 0 and ranges[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert all(0 < hi - lo <= block for lo, hi in ranges)
    for g in range(world):
        assert all((lo // block) % world == g for lo, hi in shard_blocks(N, g, world, block))


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_gather_without_process_group():
    assert gather_partials((1.5, 2.5e-17)) == [1.5, 2.5e-17]


WORKER = r'''
import ctypes, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hvcases, moocore_b200 as mb
from moocore_b200.sharding import gather_partials, shard_blocks, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
O = ctypes.CDLL(os.path.join(ROOT, "oracle", "libhvoracle.so"))
O.orc_hw_window.restype = ctypes.c_double
O.orc_hv_approx_hua_wang.restype = ctypes.c_double
dp = ctypes.POINTER(ctypes.c_double)
d, N = 4, 30000
x = hvcases.front("sphere", 64, d, 11)
tp = np.ascontiguousarray(1.1 - x)
i0, i1 = shard_range(N, rank, world)
part = O.orc_hw_window(tp.ctypes.data_as(dp), ctypes.c_size_t(64), d, ctypes.c_uint64(N), ctypes.c_uint64(i0), ctypes.c_uint64(i1 - i0), None)
flat = gather_partials((part, 0.0))
assert len(flat) == 2 * world
est = mb.hv_approx_finalize(d, N, flat)
ref = np.full(d, 1.1); mx = np.zeros(d, np.uint8)
full = O.orc_hv_approx_hua_wang(x.ctypes.data_as(dp), ctypes.c_size_t(64), ctypes.c_uint8(d), ref.ctypes.data_as(dp), mx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_uint64(N))
assert abs(est - full) / full < 1e-13, (est, full)
# block-cyclic shards (DZ2019-HW on the GPU path): same estimate
part = sum(O.orc_hw_window(tp.ctypes.data_as(dp), ctypes.c_size_t(64), d, ctypes.c_uint64(N), ctypes.c_uint64(a), ctypes.c_uint64(b - a), None)
           for a, b in shard_blocks(N, rank, world, block=4096))
est2 = mb.hv_approx_finalize(d, N, gather_partials((part, 0.0)))
assert abs(est2 - full) / full < 1e-13, (est2, full)
ests = [None] * world
dist.all_gather_object(ests, est)
assert all(e == ests[0] for e in ests)          # every rank finalises to the same bits
if rank == 0:
    print("OK", est, full)
dist.destroy_process_group()
'''


def test_two_rank_gloo_exchange(oracle, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "OK" in res.stdout


def test_shard_block_matches_the_library():
    import ctypes
    from moocore_b200.sharding import shard_block
    L = ctypes.CDLL(os.path.join(ROOT, "moocore_b200", "libmoocore_b200.so"))
    L.hvb200_shard_block.restype = ctypes.c_uint64
    for N in (1, 1000, 262144, 10**6, 10**8, 2**31):
        for world in (1, 2, 8, 16):
            assert shard_block(N, world) == L.hvb200_shard_block(ctypes.c_uint64(N), world)
            assert (1 << 10) <= shard_block(N, world) <= (1 << 14)
