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
