#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/oracle.cpp).

These fixtures are generated from the oracle (which is itself pinned against the reference's own code: see
make_reference_golden.py / reference_films.npz and DESIGN.md §2): a change of the oracle's behaviour shows up as a diff here, and
the GPU tests get fixed, machine-independent vectors (ray batches, Philox-replay films, BSDF tables).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nanogi_b200 import capi, scenes  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import parity_common as pc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ("cornell_box", "cornell_spheres"):
        sd = scenes.to_scene_data(getattr(scenes, name)(), 1.0)
        orc = pyoracle.OracleScene(sd)
        rays = np.concatenate([scenes.camera_rays(sd, 16, 16), scenes.random_rays(sd, 768, 3)])
        occ = scenes.random_rays(sd, 512, 4, occlusion=True)
        hits = orc.trace(rays, 0)
        brute = orc.trace(rays, 2)
        assert np.array_equal(hits, brute), "oracle BVH differs from its own brute force"
        occ_hits = orc.trace(occ, 1)
        # films: the scene scaled to unit size. At Cornell scale (coordinates ~550, absolute epsilon 1e-4 ~ 1.6 fp32 ulp)
        # whether a grazing ray re-hits its own surface depends on the last bit of the shading arithmetic, so fp32 device
        # code and the fp64 oracle agree only statistically there (tests/test_gpu_parity.py); at unit scale sample by sample.
        from tests.conftest import scaled_spec
        orc_s = pyoracle.OracleScene(scenes.to_scene_data(scaled_spec(getattr(scenes, name)(), 0.01), 1.0))
        films = {}
        for renderer in ("pt", "ptdirect"):
            for m in (-1, 4):
                f, st = orc_s.render(renderer, 4096, 16, 16, max_num_vertices=m, seed=3, rng_mode=1, num_threads=1)
                films[f"film_{renderer}_{m}"] = f
                films[f"rays_{renderer}_{m}"] = np.array([st["extend_rays"], st["shadow_rays"]])
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), rays=rays, hits=hits, occ=occ, occ_tri=occ_hits["tri"], **films)
        print(name, "hit rate", float((hits["tri"] != capi.NO_HIT).mean()), "occluded", float((occ_hits["tri"] == 0).mean()))
    # light tracing / raw sensor / mixed lights: unit-scale films of all four renderers
    from tests.conftest import scaled_spec
    films = {}
    for scene in ("cornell_raw_sensor", "cornell_mixed_lights"):
        spec = scenes.cornell_raw_sensor(spheres=True) if scene == "cornell_raw_sensor" else scenes.cornell_mixed_lights()
        orc_s = pyoracle.OracleScene(scenes.to_scene_data(scaled_spec(spec, 0.01), 1.0))
        for renderer in ("pt", "ptdirect", "lt", "ltdirect", "bdpt"):
            f, st = orc_s.render(renderer, 4096, 16, 16, max_num_vertices=5, seed=3, rng_mode=1, num_threads=1)
            films[f"film_{scene}_{renderer}"] = f
            films[f"rays_{scene}_{renderer}"] = np.array([st["extend_rays"], st["shadow_rays"]])
            print(scene, renderer, "mean", float(f.mean()))
    np.savez_compressed(os.path.join(OUT, "light_tracing.npz"), **films)
    # BSDF tables on the C2 scene: D wall, G conductor sphere, S fresnel sphere
    sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
    orc = pyoracle.OracleScene(sd)
    tab = {}
    for prim, bit in pc.bsdf_test_prims(sd):
        q = pc.bsdf_queries(sd, prim, bit, 256, seed=17)
        wo, fs, pdf, ok = pc.oracle_bsdf_table(orc, q)
        tab[f"q_{prim}_{bit}"] = q; tab[f"wo_{prim}_{bit}"] = wo; tab[f"fs_{prim}_{bit}"] = fs; tab[f"pdf_{prim}_{bit}"] = pdf; tab[f"ok_{prim}_{bit}"] = ok
    np.savez_compressed(os.path.join(OUT, "bsdf_tables.npz"), **tab)
    print("bsdf tables", sorted(k for k in tab if k.startswith("q_")))


if __name__ == "__main__":
    main()
