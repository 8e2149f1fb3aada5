"""GPU box: one k_bdpt launch on C2's scene for an ncu capture (see tools/ncu_summary.py blocks)."""
import sys
sys.path.insert(0, ".")
from nanogi_b200 import capi, scenes
sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
g = capi.GpuScene(sd, 0)
g.render("bdpt", 1 << 21, 1024, 1024, seed=1)
