#!/usr/bin/env python
"""BVH8 build quality on the CPU simulator (tests/hostsim: the same per-item build / traversal functions as the CUDA kernels).

    python tools/bvh_quality.py [--scene c3|c2|c4small] [--samples N]

Prints nodes, depth and the traversal cost (node steps, triangle tests per ray) of camera rays, random rays and — what the
renders actually trace — the extend and shadow rays of a small `ptdirect` render. This is where build changes are evaluated
before GPU time is spent on them.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nanogi_b200 import scenes  # noqa: E402
from tests.hostsim import pysim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c3")
    ap.add_argument("--samples", type=int, default=1 << 15)
    args = ap.parse_args()
    gen = {"c3": scenes.instanced_spheres, "c2": scenes.cornell_spheres, "c1": scenes.cornell_box,
           "c4small": lambda: scenes.interior(target_tris=1_000_000)}[args.scene]
    W, H = (1920, 1080) if args.scene in ("c3", "c4small") else (1024, 1024)
    sd = scenes.to_scene_data(gen(), W / H)
    t0 = time.time()
    sim = pysim.SimScene(sd)
    info = sim.info()
    print(f"scene {args.scene}: {info['n']} tris, {info['nodes8']} BVH8 nodes ({info['n'] / info['nodes8']:.2f} tris/node), depth {info['depth8']}, "
          f"build {time.time() - t0:.1f} s (CPU simulator)")
    cam = scenes.camera_rays(sd, 96, 54)
    rnd = scenes.random_rays(sd, 1 << 13, 3)
    print("camera rays  : node steps %.2f  tri tests %.2f" % sim.trace_stats(cam))
    print("random rays  : node steps %.2f  tri tests %.2f" % sim.trace_stats(rnd))
    _, st = sim.render("ptdirect", args.samples, W, H, seed=7, wave_capacity=8192)
    print("render ptdirect (%d samples): extend rays/path %.2f shadow rays/path %.2f" % (args.samples, st["extend_rays"] / args.samples, st["shadow_rays"] / args.samples))
    print("extend rays  : node steps %.2f  tri tests %.2f" % (st["extend_node_steps"] / st["extend_rays"], st["extend_tri_tests"] / st["extend_rays"]))
    print("shadow rays  : node steps %.2f  tri tests %.2f" % (st["shadow_node_steps"] / max(st["shadow_rays"], 1), st["shadow_tri_tests"] / max(st["shadow_rays"], 1)))


if __name__ == "__main__":
    main()
