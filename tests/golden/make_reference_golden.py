#!/usr/bin/env python
"""Generates tests/golden/reference_films.npz with THE REFERENCE'S OWN CODE (oracle/_ref/libnanogi_ref.so, built by
oracle/build_ref.sh from /root/reference/src/nanogi.cpp + include/nanogi/*.hpp against the stand-in libraries of oracle/refshim).

Unlike the other fixtures of this directory these ARE reference outputs: films of Renderer::Render (one thread, release-mode
seed std::time(nullptr) interposed) for all four GPU-path renderers on six scenes, plus tables of Primitive::SampleDirection /
EvaluateDirection / EvaluateDirectionPDF. tests/test_golden.py checks that the oracle reproduces them (mt19937 mode), on any
machine, with or without oracle/_ref. Run from the repo root in a container that has /root/reference:
    bash oracle/build_ref.sh && python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nanogi_b200 import scenes  # noqa: E402
from oracle import pyref  # noqa: E402
from tests import parity_common as pc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SCENES = {
    "cornell_box": (scenes.cornell_box, 8),
    "cornell_spheres": (scenes.cornell_spheres, -1),
    "cornell_mixed_lights": (scenes.cornell_mixed_lights, -1),
    "cornell_raw_sensor": (lambda: scenes.cornell_raw_sensor(spheres=True), 7),
    "cornell_textured": (scenes.cornell_textured, -1),
    "furnace": (lambda: scenes.furnace(0.5, 1.0), 5),
}
N, W, H, SEED = 20000, 20, 12, 20261017


def main():
    out = {"meta": np.array([N, W, H, SEED])}
    for name, (make, m) in SCENES.items():
        ref = pyref.RefScene(make(), W / H)
        for renderer in ("pt", "ptdirect", "lt", "ltdirect", "bdpt"):
            f = ref.render(renderer, N, W, H, max_num_vertices=m, seed=SEED, num_threads=1)
            out[f"film_{name}_{renderer}"] = f
            print(name, renderer, "mean", float(f.mean()))
        ref.close()
    # Primitive function tables on the C2 materials, both transport directions
    spec = scenes.cornell_spheres()
    sd = scenes.to_scene_data(spec, 1.0)
    ref = pyref.RefScene(spec, 1.0)
    for prim, bit in pc.BSDF_TEST_PRIMS:
        q = pc.bsdf_queries(sd, prim, bit, 256, seed=23)
        wo = np.zeros((q.shape[0], 3)); fs = np.zeros((q.shape[0], 2, 3)); pdf = np.zeros((q.shape[0], 2))
        for i in range(q.shape[0]):
            sn, gn, wi = q[i, 2:5].astype(np.float64), q[i, 5:8].astype(np.float64), q[i, 8:11].astype(np.float64)
            wo[i] = ref.sample_direction(prim, bit, sn, gn, wi, float(q[i, 11]), float(q[i, 12]), float(q[i, 13]))
            for k, el in enumerate((True, False)):
                fs[i, k], pdf[i, k] = ref.evaluate_direction(prim, bit, sn, gn, wi, wo[i], el, True)
        out[f"q_{prim}_{bit}"] = q; out[f"wo_{prim}_{bit}"] = wo; out[f"fs_{prim}_{bit}"] = fs; out[f"pdf_{prim}_{bit}"] = pdf
    ref.close()
    np.savez_compressed(os.path.join(OUT, "reference_films.npz"), **out)
    print("wrote reference_films.npz")


if __name__ == "__main__":
    main()
