"""Sample-index sharding of one render across ranks.

The reference runs independent chunks of sample indices with private films and one final sum
(src/nanogi.cpp:281-337, :429-437). Here rank r of G takes the contiguous range [r N / G, (r+1) N / G);
with counter-based Philox keyed by the sample index the SET of samples is identical for any G, only the
summation order (hence the last ulp) changes. Every rank pre-scales its splats by W*H/N, so the film is
the plain sum of the per-rank films: one reduce(SUM) to rank 0 over NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(num_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(sample_offset, num_samples) of `rank`."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    lo = num_samples * rank // world_size
    hi = num_samples * (rank + 1) // world_size
    return lo, hi - lo


def reduce_film(film, dst: int = 0):
    """Sums the per-rank films onto `dst` (torch tensor, in place). No-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
