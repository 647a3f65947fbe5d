"""Multi-GPU sharding of one hv_approx estimate.

The N directions are independent, so rank g of G evaluates its share — block-cyclic blocks for DZ2019-HW
(shard_blocks), the contiguous index range [g*N/G, (g+1)*N/G) for the sequential MC and Rphi streams
(shard_range) — on its own GPU and produces one (hi, lo) FP64 pair; the only exchange step
is an all-gather of those 16 bytes per rank (NCCL over NVLink when the process group is
"nccl", gloo in the CPU tests), followed by a fixed-order long double sum and the c_d/N
scaling of c/hvapprox.c:691-692.  An all-gather + ordered sum is used instead of an all-reduce
so that the result does not depend on the collective's reduction order.
"""
from __future__ import annotations

from typing import Sequence, Tuple


def shard_range(nsamples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping index range of `rank`."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    return (nsamples * rank) // world, (nsamples * (rank + 1)) // world


SHARD_BLOCK = 1 << 14      # HVB200_SHARD_BLOCK of include/hvapprox_b200.h


def shard_block(nsamples: int, world: int) -> int:
    """Block size used for `world` ranks (hvb200_shard_block): 2^14 directions, halved down to 2^10 until every
    rank gets at least 4 blocks."""
    block = SHARD_BLOCK
    while block > (1 << 10) and block * max(1, world) * 4 > nsamples:
        block >>= 1
    return block


def shard_blocks(nsamples: int, rank: int, world: int, block: int = SHARD_BLOCK):
    """Block-cyclic shard of `rank`: the index ranges [b*block, min((b+1)*block, nsamples)) for b = rank,
    rank + world, ...  — what hvb200_plan_run_blocks evaluates for (first=rank, stride=world).  Used for
    DZ2019-HW, whose per-direction cost drifts along the lattice index."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    if block <= 0:
        raise ValueError("block must be positive")
    return [(lo, min(lo + block, nsamples)) for lo in range(rank * block, nsamples, world * block)]


def mc_shard_start(plan, nsamples: int, seed: int, rank: int, world: int, group=None, device=None):
    """Sample range [i0, i1) of `rank` for a DZ2019-MC estimate, with the plan primed to start its parse there.

    The MT19937/ziggurat stream is sequential (c/rng.c:71-99), so a contiguous sample range starting at i0 > 0 costs a
    parse of the whole prefix (rank 7 of 8 parsed 7/8 of the stream).  Here every rank describes its own PAIR range of
    the stream (Plan.mc_segment_scan: one pass over 1/world of it), the descriptions (21 words) are all-gathered, and
    every rank resolves the same partition of [0, nsamples) from them (mc_segments_resolve); per-sample values are the
    sequential stream's bit for bit.  Short streams, HVB200_MC=host and HVB200_MC_SHARD=prefix keep shard_range."""
    import os

    from . import _api

    lay = None
    if not plan.empty and os.environ.get("HVB200_MC") != "host" and os.environ.get("HVB200_MC_SHARD") != "prefix":
        lay = _api.mc_segment_layout(nsamples, plan.nobj, world, rank)
    if lay is None:
        return shard_range(nsamples, rank, world)
    try:
        mine = plan.mc_segment_scan(seed, lay[0], lay[1]).to_list()
    except RuntimeError:
        # a tail loop beyond the device tables or a near-tie decision in this rank's range (~1e-12 per pair): every rank
        # still takes part in the exchange, sees the marker and falls back to sample ranges together
        mine = [lay[0], lay[1], _api.MC_OPEN_END] + [0] * 18
    words = gather_words(mine, group=group, device=device)
    segs = [_api.McSegment.from_list(words[g * _api.McSegment.WORDS:(g + 1) * _api.McSegment.WORDS]) for g in range(world)]
    if any(sg.merge == _api.MC_OPEN_END for sg in segs):
        return shard_range(nsamples, rank, world)
    pair, discard, i0, i1 = _api.mc_segments_resolve(segs, plan.nobj, nsamples)[rank]
    plan.set_mc_start(pair, discard)
    return i0, i1


def gather_words(words: Sequence[int], group=None, device=None):
    """All-gather a short list of unsigned 64-bit words per rank (as int64 bit patterns); flat list, rank-major."""
    import torch
    import torch.distributed as dist

    def wrap(v):
        v = int(v)
        return v - (1 << 64) if v >= (1 << 63) else v

    if not (dist.is_available() and dist.is_initialized()):
        return [int(w) for w in words]
    world = dist.get_world_size(group)
    if dist.get_backend(group) != "nccl":
        device = None                            # gloo (CPU tests): host tensors
    elif isinstance(device, int):
        device = torch.device("cuda", device)
    send = torch.tensor([wrap(w) for w in words], dtype=torch.int64, device=device)
    recv = torch.empty(len(words) * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return [v + (1 << 64) if v < 0 else v for v in recv.cpu().tolist()]


_gather_cache = {}


def gather_partials(hilo: Sequence[float], group=None, device=None):
    """All-gather one (hi, lo) pair per rank; returns a flat list [hi0, lo0, hi1, lo1, ...].

    Works with any initialised torch.distributed backend; without an initialised process
    group it returns the local pair (world size 1).  The send/receive tensors (and, on a GPU, a pinned
    host landing buffer) are allocated once per (device, world) and reused: one collective and one
    device->host copy per call."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [float(hilo[0]), float(hilo[1])]
    world = dist.get_world_size(group)
    key = (str(device), world, id(group))
    bufs = _gather_cache.get(key)
    if bufs is None:
        send = torch.empty(2, dtype=torch.float64, device=device)
        recv = torch.empty(2 * world, dtype=torch.float64, device=device)
        on_gpu = send.is_cuda
        stage = torch.empty(2, dtype=torch.float64, pin_memory=on_gpu)
        land = torch.empty(2 * world, dtype=torch.float64, pin_memory=on_gpu)
        bufs = _gather_cache[key] = (send, recv, stage, land)
    send, recv, stage, land = bufs
    stage[0] = float(hilo[0])
    stage[1] = float(hilo[1])
    send.copy_(stage, non_blocking=True)
    dist.all_gather_into_tensor(recv, send, group=group)
    land.copy_(recv, non_blocking=True)
    if recv.is_cuda:
        torch.cuda.current_stream(recv.device).synchronize()
    return land.tolist()
