import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # measured numbers: profiles/r02_sim_vs_gpu_sfu.txt
import numpy as np
from tests.hostsim import pysim
from nanogi_b200 import scenes, capi
from tests.test_gpu_parity import scaled_spec
small = scenes.to_scene_data(scaled_spec(scenes.cornell_box(), 0.01), 1.0)
cornell = scenes.to_scene_data(scenes.cornell_box(), 1.0)
for name, sd in (("small", small), ("cornell", cornell)):
    g = capi.GpuScene(sd, 0); sim = pysim.SimScene(sd)
    for r in ("pt", "ptdirect"):
        for seed in (12, 13):
            fg, sg = g.render(r, 30000, 32, 32, seed=seed, max_num_vertices=8)
            fs, ss = sim.render(r, 30000, 32, 32, seed=seed, max_num_vertices=8)
            close = np.isclose(fg, fs, rtol=2e-3, atol=1e-5 * fs.max()).all(axis=2)
            print(name, r, seed, "extend", sg.extend_rays, ss["extend_rays"], "shadow", sg.shadow_rays, ss["shadow_rays"], "bad pixels", float((~close).mean()), "means", float(fg.mean()), float(fs.mean()), flush=True)
    g.close()
