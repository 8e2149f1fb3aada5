"""Multi-rank host logic on CPU: world_size-2 gloo process group. What runs here is what `bench.py --gpus N` runs
over NCCL: nanogi_b200.shard.shard_range (-> ngi_gpu_shard_range in libnanogi_gpu.so, the arithmetic ngi_gpu_group_render
uses too), shard.exchange_comm_id (rank 0's NCCL id, made by ngi_gpu_comm_get_id, to every rank) and shard.reduce_film (one
reduce(SUM) of the per-rank films to rank 0; over gloo here, through the product's ncclReduce when a communicator exists).
The per-rank renderer here is the CPU oracle in Philox mode (test infrastructure), so the sharded film must equal
the single-process film of the same sample set up to fp64 summation order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

N, W, H, SEED = 20000, 24, 24, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, renderer, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanogi_b200 import scenes, shard
    from oracle import pyoracle
    sd = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    orc = pyoracle.OracleScene(sd)
    off, cnt = shard.shard_range(N, rank, world)
    # the communicator id travels like in bench.py: made by the product on rank 0, broadcast over the process group
    import ctypes
    from nanogi_b200 import capi
    blob = None
    if rank == 0:
        cid = capi.NgiCommId()
        rc = capi.gpu_lib().ngi_gpu_comm_get_id(ctypes.byref(cid))
        blob = ctypes.string_at(ctypes.byref(cid), 128) if rc == 0 else bytes(range(128))     # no libnccl on this host: any 128 bytes
    got = shard.exchange_comm_id(blob)
    ids = [None, None]
    dist.all_gather_object(ids, got)
    assert ids[0] == ids[1] and len(got) == 128 and (rank != 0 or got == blob)
    film, st = orc.render(renderer, cnt, W, H, max_num_vertices=6, seed=SEED, rng_mode=1, num_threads=1,
                          sample_offset=off, film_norm_samples=N)
    t = torch.from_numpy(film)
    shard.reduce_film(t, dst=0)
    rays = torch.tensor([st["extend_rays"], st["shadow_rays"]], dtype=torch.float64)
    dist.reduce(rays, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(out_path, film=t.numpy(), rays=rays.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from nanogi_b200 import shard
    for n in (0, 1, 7, 1000, 2**31 + 12345, 8493465600):
        for g in (1, 2, 3, 8):
            parts = [shard.shard_range(n, r, g) for r in range(g)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect", "bdpt"])
def test_two_rank_gloo_render_equals_single_process(renderer, tmp_path):
    out = str(tmp_path / "film.npz")
    mp.spawn(_worker, args=(2, _free_port(), renderer, out), nprocs=2, join=True)
    got = np.load(out)
    sys.path.insert(0, ROOT)
    from nanogi_b200 import scenes
    from oracle import pyoracle
    sd = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    ref, st = pyoracle.OracleScene(sd).render(renderer, N, W, H, max_num_vertices=6, seed=SEED, rng_mode=1, num_threads=1)
    assert got["rays"][0] == st["extend_rays"] and got["rays"][1] == st["shadow_rays"]
    assert np.allclose(got["film"], ref, rtol=1e-12, atol=1e-15)
