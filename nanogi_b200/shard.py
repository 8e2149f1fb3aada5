"""Sample-index sharding of one render across ranks — the host logic `bench.py --gpus N` runs.

The reference runs independent chunks of sample indices with private films and one final sum
(src/nanogi.cpp:281-337, :429-437). Here rank r of G takes the contiguous range [r N / G, (r+1) N / G)
(`ngi_gpu_shard_range`, the product's own arithmetic: the in-process `ngi_gpu_group_render` uses the same function);
with counter-based Philox keyed by the sample index the SET of samples is identical for any G, only the
summation order (hence the last ulp) changes. Every rank pre-scales its splats by W*H/N, so the film is
the plain sum of the per-rank films: one reduce(SUM) to rank 0 — through the product's NCCL communicator
(`capi.GpuComm`, ncclReduce over NVLink) when one is passed, else through the torch.distributed process group
(gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(num_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(sample_offset, num_samples) of `rank`."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    from . import capi
    return capi.shard_range(num_samples, rank, world_size)


def exchange_comm_id(blob: Optional[bytes], src: int = 0) -> bytes:
    """Hands rank `src`'s 128-byte NCCL id to every rank over the torch.distributed process group (any backend)."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        assert blob is not None and len(blob) == 128
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def make_comm(device: int):
    """The product's NCCL communicator for this rank (one process per GPU), or None for a single rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    from . import capi
    return capi.GpuComm(dist.get_rank(), dist.get_world_size(), device, exchange_comm_id)


def reduce_film(film, dst: int = 0, comm=None, stream_ptr: int = 0):
    """Sums the per-rank films onto `dst` (torch tensor, in place). No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return film
    if comm is not None:
        comm.reduce_film(film.data_ptr(), film.numel(), dst, stream_ptr)      # ncclReduce on the render's stream
    else:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
