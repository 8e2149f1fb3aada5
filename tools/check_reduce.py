#!/usr/bin/env python
"""N ranks (torchrun): the product's film reduce (ngi_gpu_comm_reduce_film, in place) against known patterns and against torch.distributed.reduce."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from nanogi_b200 import shard
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = shard.make_comm(local)
    stream = torch.cuda.current_stream(dev)
    n = 1920 * 1080 * 3
    res = []
    for it in range(4):
        g = torch.Generator(device=dev); g.manual_seed(1000 * it + rank)
        film = torch.rand(n, dtype=torch.float32, device=dev, generator=g)
        film[(rank * 7919 + it) % n] = 1e6 * (rank + 1)            # a "firefly" per rank
        ref = film.clone()
        dist.reduce(ref, dst=0, op=dist.ReduceOp.SUM)
        mine = film.clone()
        shard.reduce_film(mine, 0, comm, stream.cuda_stream)
        torch.cuda.synchronize()
        if rank == 0:
            res.append({"iter": it, "max_abs_diff_vs_torch": float((mine - ref).abs().max()), "sum_mine": float(mine.double().sum()), "sum_torch": float(ref.double().sum())})
        else:
            # the send buffer of a non-root rank must come back untouched
            res.append(bool(torch.equal(mine, film)))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        print(json.dumps({"world": world, "root": gathered[0], "non_root_send_buffers_untouched": [all(x) for x in gathered[1:]]}))
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
