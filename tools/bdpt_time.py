import sys; sys.path.insert(0, ".")
from nanogi_b200 import capi, scenes
sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
g = capi.GpuScene(sd, 0)
g.render("bdpt", 1 << 20, 1024, 1024, seed=1)
per_thread = "--per-thread" in sys.argv
for m in (2, 3, 4, 6, 8, 12, -1):
    n = 1 << 24
    f, st = g.render("bdpt", n, 1024, 1024, max_num_vertices=m, seed=2, flags=capi.RENDER_BDPT_PER_THREAD if per_thread else 0)
    print("bdpt", "per-thread" if per_thread else "wavefront", "m", m, "Mpaths/s %.1f Mrays/s %.1f mean %.5g launches %d" % (n / st.gpu_seconds / 1e6, (st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6, f.mean(), st.kernel_launches), flush=True)
