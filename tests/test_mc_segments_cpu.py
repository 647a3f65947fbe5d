"""Host logic of the DZ2019-MC pair-range shards (hvb200_mc_segment_layout / hvb200_mc_segments_resolve,
moocore_b200/csrc/hvapprox_b200.c): the stream of c/rng.c:71-99 is modelled as a random sequence of attempts — an attempt
that starts at pair j consumes L[j] pairs and emits E[j] normals (0 or 1) — exactly what the device classification
(k_zig_classify) produces, and the resolved starts are checked against the sequential parse."""
import os
import subprocess
import sys

import numpy as np
import pytest

import moocore_b200 as mb
from moocore_b200._api import MC_OPEN_END, McSegment, mc_segment_layout, mc_segments_resolve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model_stream(npairs, rng, p_long=0.03):
    L = np.ones(npairs, dtype=np.int64)
    long_ = rng.random(npairs) < p_long
    L[long_] = rng.choice([2, 2, 2, 3, 5, 7, 16], size=int(long_.sum()))
    E = (rng.random(npairs) < 0.99).astype(np.int64)
    return L, E


def sequential_parse(L, E):
    """start[j] = True where the parse from pair 0 starts an attempt; before[j] = normals emitted before pair j."""
    n = len(L)
    start = np.zeros(n, dtype=bool)
    before = np.zeros(n + 17, dtype=np.int64)
    j = q = 0
    while j < n:
        start[j] = True
        nxt = j + L[j]
        before[j:nxt] = q          # pairs inside the attempt: the count in front of the attempt
        q += E[j]
        j = nxt
    before[n:] = q
    return start, before


def describe(L, E, pair0, npairs):
    """What mcgen_scan_segment computes on the device, by brute force."""
    end = len(L) if npairs == MC_OPEN_END else pair0 + npairs
    pos = [pair0 + s for s in range(16)]
    cnt = [0] * 16
    while len(set(pos)) > 1:
        p = min(pos)
        for s in range(16):
            if pos[s] == p:
                pos[s] = p + L[p]
                cnt[s] += E[p]
    merge = pos[0] - pair0
    body = 0
    skip_out = 0
    if npairs != MC_OPEN_END:
        j = pos[0]
        while j < end:
            body += E[j]
            j += L[j]
        skip_out = j - end
    return McSegment(pair0, npairs, merge, (mb._api.ctypes.c_uint64 * 16)(*cnt), body, skip_out)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("d", [1, 3, 50])
def test_resolve_matches_the_sequential_parse(world, d):
    rng = np.random.default_rng(100 * world + d)
    npairs = 60000
    L, E = model_stream(npairs, rng)
    start, before = sequential_parse(L, E)
    total = int(before[npairs])
    N = (total - 200) // d                       # the stream holds a little more than N samples
    for trial in range(5):
        cuts = np.sort(rng.choice(np.arange(2000, npairs - 5000), size=world - 1, replace=False))
        if np.any(np.diff(np.concatenate([[0], cuts])) < 200):
            continue
        bounds = [0, *cuts.tolist()]
        segs = [describe(L, E, bounds[g], (bounds[g + 1] - bounds[g]) if g + 1 < world else MC_OPEN_END) for g in range(world)]
        starts = mc_segments_resolve(segs, d, N)
        assert starts[0][2] == 0 and starts[-1][3] == N
        for g, (pair, discard, i0, i1) in enumerate(starts):
            assert start[pair], "a rank must begin at a real attempt start of the sequential parse"
            if i0 < N:
                assert before[pair] + discard == i0 * d          # the first kept normal opens sample i0
                assert discard < d
            if g + 1 < world:
                assert i1 == starts[g + 1][2]


def test_resolve_rejects_gaps_and_bad_heads():
    rng = np.random.default_rng(5)
    L, E = model_stream(30000, rng)
    a = describe(L, E, 0, 10000)
    b = describe(L, E, 10001, MC_OPEN_END)       # not adjacent
    with pytest.raises(RuntimeError):
        mc_segments_resolve([a, b], 3, 1000)
    b = describe(L, E, 10000, MC_OPEN_END)
    a.skip_out = 16
    with pytest.raises(RuntimeError):
        mc_segments_resolve([a, b], 3, 1000)


def test_layout_partitions_the_expected_stream():
    # short streams keep contiguous sample ranges
    assert mc_segment_layout(1000, 5, 4, 0) is None
    assert mc_segment_layout(10**6, 5, 2, 0) is None                 # cfg2: 2.5e6 pairs per rank, below the 2^26 threshold
    assert mc_segment_layout(10**9, 50, 1, 0) is None
    N, d, world = 10**9, 50, 8
    lay = [mc_segment_layout(N, d, world, g) for g in range(world)]
    assert lay[0][0] == 0 and lay[-1][1] == MC_OPEN_END
    for g in range(world - 1):
        assert lay[g][0] + lay[g][1] == lay[g + 1][0]
    # ~1.022 pairs per normal: the open-ended last range gets about the same share as the others
    W = lay[0][1]
    assert abs(W * world / (N * d) - 1.02201) < 1e-4
    with pytest.raises(ValueError):
        mc_segment_layout(N, d, 4, 4)


WORKER = r'''
import os, sys
import torch.distributed as dist
sys.path.insert(0, ROOT)
from moocore_b200.sharding import gather_words
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = [rank, (1 << 64) - 1 - rank, 1 << 63, 12345678901234567 + rank]
flat = gather_words(mine)
assert len(flat) == 4 * world
for g in range(world):
    assert flat[4 * g:4 * g + 4] == [g, (1 << 64) - 1 - g, 1 << 63, 12345678901234567 + g], flat
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def test_gather_words_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29950 + os.getpid() % 40), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "OK" in res.stdout


WORKER_FALLBACK = r'''
import os, sys
import torch.distributed as dist
sys.path.insert(0, ROOT)
from moocore_b200 import _api
from moocore_b200.sharding import mc_shard_start, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, d = 10**9, 50

class FakePlan:                     # the exchange protocol without a GPU: rank 1's scan fails
    empty = False
    nobj = d
    def mc_segment_scan(self, seed, pair0, npairs):
        if rank == 1:
            raise RuntimeError("forced: tail loop beyond the device tables")
        seg = _api.McSegment(pair0, npairs, 15, (_api.ctypes.c_uint64 * 16)(*([0] * 16)), 1000, 0)
        return seg
    def set_mc_start(self, pair, discard):
        raise AssertionError("no rank may start from a pair range when one scan failed")

assert _api.mc_segment_layout(N, d, world, rank) is not None
got = mc_shard_start(FakePlan(), N, 42, rank, world)
assert got == shard_range(N, rank, world), got
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def test_failed_scan_sends_every_rank_to_sample_ranges(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER_FALLBACK)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 40), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "OK" in res.stdout
