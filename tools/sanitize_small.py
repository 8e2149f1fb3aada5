#!/usr/bin/env python
"""Small end-to-end exercise of the CUDA module for compute-sanitizer (memcheck / racecheck / initcheck), run on the GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py      (shared-memory pool of the trace loop, block_reserve)

Scene build (PLOC + SAH-optimal collapse + expand), ray queries (BVH8 / BVH2 / brute force), pt / ptdirect / ltdirect / bdpt renders
with small waves (several lanes, ramp-up and drain), a one-device group render and an accumulate render — all checked against the oracle
where that is cheap."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanogi_b200 import capi, scenes  # noqa: E402
from oracle import pyoracle  # noqa: E402


def quick():
    """`--quick`: the shared-memory users only (trace pool of k_trace8 / k_extend / k_shadow / k_bdw_*, block_reserve of the logic kernels) on the
    smallest workload that reaches every phase — racecheck instruments every shared access and needs ~1 s per persistent launch; run it with
    NGI_TRACE_GRID_PCT=6 NGI_LANES=1 so that the trace grids are one CTA per SM."""
    sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
    g = capi.GpuScene(sd, 0)
    orc = pyoracle.OracleScene(sd)
    rays = scenes.random_rays(sd, 3000, 3)
    ho, hg = orc.trace(rays, 0), g.trace(rays, False, 0)
    assert np.array_equal(hg["tri"], ho["tri"]) and np.array_equal(hg["t"], ho["t"])
    occ = scenes.random_rays(sd, 3000, 4, occlusion=True)
    assert np.array_equal(g.trace(occ, True, 0)["tri"], orc.trace(occ, 1)["tri"])
    print("trace ok", flush=True)
    f, st = g.render("ptdirect", 3000, 16, 16, max_num_vertices=5, seed=3, wave_capacity=1024)
    assert np.isfinite(f).all() and st.paths == 3000
    print("ptdirect ok", flush=True)
    f, st = g.render("bdpt", 400, 16, 16, max_num_vertices=4, seed=3)
    assert np.isfinite(f).all() and st.paths == 400
    print("bdpt ok", flush=True)
    g.close()
    print("sanitize_small quick OK")


def main():
    if "--quick" in sys.argv:
        return quick()
    for name, gen in (("cornell_spheres", scenes.cornell_spheres), ("cornell_branches", lambda: scenes.cornell_branches(light_res=8))):
        sd = scenes.to_scene_data(gen(), 1.0)
        g = capi.GpuScene(sd, 0)
        orc = pyoracle.OracleScene(sd)
        rays = np.concatenate([scenes.camera_rays(sd, 48, 48), scenes.random_rays(sd, 6000, 3)])
        ho = orc.trace(rays, 0)
        for accel in (0, 1, 2):
            hg = g.trace(rays, False, accel)
            assert np.array_equal(hg["tri"], ho["tri"]) and np.array_equal(hg["t"], ho["t"]), (name, accel)
        occ = scenes.random_rays(sd, 6000, 4, occlusion=True)
        assert np.array_equal(g.trace(occ, True, 0)["tri"], orc.trace(occ, 1)["tri"])
        for renderer, n in (("pt", 40000), ("ptdirect", 40000), ("ltdirect", 20000), ("bdpt", 8000)):
            f, st = g.render(renderer, n, 32, 32, max_num_vertices=6, seed=3, wave_capacity=4096)
            assert np.isfinite(f).all() and st.paths == n
        f1, _ = g.render("ptdirect", 30000, 32, 32, seed=5, wave_capacity=2048)
        f2, _ = g.render("ptdirect", 30000, 32, 32, seed=5, wave_capacity=1 << 16)
        assert np.allclose(f1, f2, rtol=1e-3, atol=1e-6 * f1.max())
        g.close()
        grp = capi.GpuGroup(sd, [0])
        fg, sg = grp.render("ptdirect", 30000, 32, 32, seed=5)
        assert np.allclose(fg, f1, rtol=1e-3, atol=1e-6 * f1.max())
        grp.close()
        print(name, "ok", flush=True)
    print("sanitize_small OK")


if __name__ == "__main__":
    main()
