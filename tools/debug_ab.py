import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanogi_b200 import capi, scenes
for name, W in (("cornell_box", 32), ("cornell_spheres", 32)):
    sd = scenes.to_scene_data(getattr(scenes, name)(), 1.0)
    g = capi.GpuScene(sd, 0)
    for renderer in ("pt", "ptdirect"):
        for n in (30000, 262144):
            a, sa = g.render(renderer, n, W, W, seed=12, max_num_vertices=8)
            b, sb = g.render(renderer, n, W, W, seed=12, max_num_vertices=8, flags=capi.RENDER_PER_RAY_TRACE)
            c, sc = g.render(renderer, n, W, W, seed=12, max_num_vertices=8)
            bad = (~np.isclose(a, b, rtol=1e-4, atol=1e-6 * b.max())).any(axis=2).sum()
            bad2 = (~np.isclose(a, c, rtol=1e-4, atol=1e-6 * b.max())).any(axis=2).sum()
            print(name, renderer, n, "ext", sa.extend_rays, sb.extend_rays, "sh", sa.shadow_rays, sb.shadow_rays, "bad px warp-vs-perray", bad, "warp-vs-warp", bad2,
                  "means", a.mean(), b.mean(), flush=True)
    g.close()
