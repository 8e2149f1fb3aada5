#!/usr/bin/env python
"""C5 — ray-throughput microbench + full-size geometry parity (BASELINE.json configs[4]).

    python tools/raybench.py [--scene c3|c4|c2] [--rays 16777216] [--check N]   (GPU box)

For the scene's BVH: a coherent batch (sqrt(rays) x sqrt(rays) pinhole grid from the scene camera) and an incoherent
batch (origins uniform in the AABB, directions uniform on the sphere, seed 3), each as closest-hit and as occlusion
(tfar = U[0.1, 1] x AABB diagonal). Device-resident rays/hits (torch tensors), kernel time from ngi_gpu_trace_device's
CUDA events, 3 warm-up + 5 timed launches. Parity: the first `--check` rays of every batch are traced by the CPU
oracle (the declared reference intersector) and compared bit for bit: primitive id, t, u, v / occlusion flag.
Prints one JSON line.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import b_ray, measured_peaks  # noqa: E402
from nanogi_b200 import capi, scenes  # noqa: E402

SCENES = {"c2": ("cornell_spheres", {}), "c3": ("instanced_spheres", {}), "c4": ("interior", {})}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c3", choices=sorted(SCENES))
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--check", type=int, default=1 << 24)
    ap.add_argument("--tris", type=int, default=0, help="c4 only: target triangle count (default 10M)")
    args = ap.parse_args()
    import torch

    gen, kw = SCENES[args.scene]
    if args.scene == "c4" and args.tris:
        kw = {"target_tris": args.tris}
    t0 = time.perf_counter()
    sd = scenes.to_scene_data(getattr(scenes, gen)(**kw), 16 / 9 if args.scene != "c2" else 1.0)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = capi.GpuScene(sd, 0)
    t_create = time.perf_counter() - t0
    info = g.info()
    side = int(math.isqrt(args.rays))
    batches = {
        "coherent": scenes.camera_rays(sd, side, side),
        "incoherent": scenes.random_rays(sd, args.rays, 3),
        "coherent_occlusion": None, "incoherent_occlusion": scenes.random_rays(sd, args.rays, 4, occlusion=True),
    }
    co = batches["coherent"].copy()
    diag = float(np.linalg.norm(sd.positions.reshape(-1, 3).max(0) - sd.positions.reshape(-1, 3).min(0)))
    co["tmax"] = (np.random.default_rng(5).uniform(0.1, 1.0, co.shape[0]) * diag).astype(np.float32)
    batches["coherent_occlusion"] = co
    peak, peak_src = measured_peaks()
    br = b_ray(int(info.num_tris))
    out = {"scene": args.scene, "tris": int(info.num_tris), "bvh8_nodes": int(info.bvh8_nodes), "bvh8_depth": int(info.bvh8_max_depth),
           "device_bytes": int(info.device_bytes), "bvh_build_ms": info.build_gpu_seconds * 1e3, "scene_create_s": t_create,
           "scene_generate_s": t_gen, "bytes_per_ray": br, "hbm_peak_gbs": peak, "peak_source": peak_src, "batches": {}}
    orc = None
    for name, rays in batches.items():
        any_hit = name.endswith("occlusion")
        n = rays.shape[0]
        d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda()
        d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        for _ in range(3):
            g.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), any_hit, 0)
        ts = [g.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), any_hit, 0) for _ in range(5)]
        sec = float(np.median(ts))
        hits = d_hits.cpu().numpy().view(capi.HIT_DTYPE).reshape(n)
        rec = {"rays": n, "ms": sec * 1e3, "grays_per_s": n / sec / 1e9, "roofline_frac_hbm": n / sec * br / 1e9 / peak,
               "hit_rate": float((hits["tri"] != capi.NO_HIT).mean())}
        if args.check > 0:
            from oracle import pyoracle
            if orc is None:
                t0 = time.perf_counter()
                orc = pyoracle.OracleScene(sd)
                out["oracle_build_s"] = time.perf_counter() - t0
            k = min(args.check, n)
            t0 = time.perf_counter()
            ho = orc.trace(rays[:k], 1 if any_hit else 0)
            rec["oracle_mrays_per_s"] = k / (time.perf_counter() - t0) / 1e6
            if any_hit:
                bad = int((hits["tri"][:k] != ho["tri"]).sum())
            else:
                bad = int(((hits["tri"][:k] != ho["tri"]) | (hits["t"][:k] != ho["t"]) | (hits["u"][:k] != ho["u"]) | (hits["v"][:k] != ho["v"])).sum())
            rec["checked"] = k
            rec["mismatches"] = bad
        out["batches"][name] = rec
        del d_rays, d_hits
    g.close()
    print(json.dumps(out), flush=True)
    if any(b.get("mismatches", 0) for b in out["batches"].values()):
        sys.exit(1)


if __name__ == "__main__":
    main()
