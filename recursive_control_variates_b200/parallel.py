"""Multi-GPU plumbing: one process per GPU, spp sharding with global lane indices.

The path shards naturally: rank g of G renders the samples
k in [g*spp/G, (g+1)*spp/G) of EVERY pixel, keeping the global lane index
pixel*spp + k for the PCG32 stream, so the union over ranks is exactly the
single-GPU sample set.  Two exchanges per iteration: all-reduce(sum) of the
film accumulators (before the weight division) and of the gradient grids.
torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is
plumbing only.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_spp(spp: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous sample sub-range [begin, end) of every pixel owned by `rank`.
    Ranges are disjoint, cover [0, spp) and differ in size by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world size")
    base, rem = divmod(spp, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sum_(tensors: Iterable[torch.Tensor]) -> None:
    """In-place sum over ranks (no-op for a single process)."""
    if world()[1] == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
