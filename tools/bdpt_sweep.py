"""GPU box: wavefront bdpt throughput against the batch size (samples per batch) on C2's scene."""
import sys; sys.path.insert(0, ".")
from nanogi_b200 import capi, scenes
sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
g = capi.GpuScene(sd, 0)
g.render("bdpt", 1 << 20, 1024, 1024, seed=1)
n = 1 << 25
for m in (-1, 6):
    for lb in (17, 18, 19, 20, 21, 22):
        g.render("bdpt", n, 1024, 1024, max_num_vertices=m, seed=2, wave_capacity=1 << lb)
        f, st = g.render("bdpt", n, 1024, 1024, max_num_vertices=m, seed=2, wave_capacity=1 << lb)
        print("bdpt m", m, "batch 2^%d" % lb, "Mpaths/s %.1f Mrays/s %.1f launches %d" % (n / st.gpu_seconds / 1e6, (st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6, st.kernel_launches), flush=True)
