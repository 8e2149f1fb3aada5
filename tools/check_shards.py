#!/usr/bin/env python
"""One GPU: the film of a job rendered as G shards by sample index (what G ranks render under bench.py --gpus G / ngi_gpu_group_render)
summed on the device equals the film of the job rendered in one piece — same Philox sample set, so means and ray counts agree.

    python tools/check_shards.py [--workload c3] [--spp 1024] [--shards 8]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nanogi_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1005)
    a = ap.parse_args()
    import torch
    gen, renderer, W, H, spp, m, desc = bench.WORKLOADS[a.workload]
    spp = a.spp or spp
    n = W * H * spp
    sd = bench.build_scene(a.workload, W / H)
    scene = capi.GpuScene(sd, 0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    full = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    st = scene.render_device(full.data_ptr(), stream.cuda_stream, renderer, n, W, H, max_num_vertices=m, seed=a.seed, film_norm_samples=n)
    torch.cuda.synchronize()
    out = {"workload": desc, "samples": n, "full": {"mean": float(full.double().mean()), "extend": st.extend_rays, "shadow": st.shadow_rays}}
    for G in sorted({a.shards, 4}):
        acc = torch.zeros((H, W, 3), dtype=torch.float64, device=dev)
        part = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        ext = sh = 0
        means = []
        for r in range(G):
            off, cnt = capi.shard_range(n, r, G)
            s = scene.render_device(part.data_ptr(), stream.cuda_stream, renderer, cnt, W, H, max_num_vertices=m, seed=a.seed, sample_offset=off, film_norm_samples=n)
            torch.cuda.synchronize()
            acc += part.double()
            means.append(float(part.double().mean()))
            ext += s.extend_rays; sh += s.shadow_rays
        d = (acc - full.double()).abs()
        out[f"shards_{G}"] = {"mean": float(acc.mean()), "extend": ext, "shadow": sh, "shard_means": means,
                              "max_abs_diff": float(d.max()), "max_rel_diff_of_image_max": float(d.max() / full.double().max()),
                              "samples_per_shard": n // G}
    print(json.dumps(out))
    scene.close()


if __name__ == "__main__":
    main()
