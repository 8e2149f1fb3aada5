"""GPU box: the wavefront bdpt on the C3 scene (1 M triangles): throughput, and its image mean next to ptdirect's / the per-thread kernel's."""
import sys; sys.path.insert(0, ".")
import numpy as np
from nanogi_b200 import capi, scenes
sd = scenes.to_scene_data(scenes.instanced_spheres(), 1920 / 1080, name="c3")
g = capi.GpuScene(sd, 0)
W, H = 960, 540
n = W * H * 64
g.render("bdpt", n // 8, W, H, max_num_vertices=6, seed=1)
for m in (6, -1):
    f, st = g.render("bdpt", n, W, H, max_num_vertices=m, seed=2)
    print("c3 bdpt wavefront m", m, "Mpaths/s %.1f Mrays/s %.1f mean %.5f finite %s" % (n / st.gpu_seconds / 1e6, (st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6, f.mean(dtype=np.float64), bool(np.isfinite(f).all())), flush=True)
f2, st2 = g.render("bdpt", n // 8, W, H, max_num_vertices=6, seed=2, flags=capi.RENDER_BDPT_PER_THREAD)
f1, st1 = g.render("bdpt", n // 8, W, H, max_num_vertices=6, seed=2)
print("c3 bdpt per-thread m 6 Mpaths/s %.1f mean %.5f; wavefront same samples mean %.5f; rays equal %s; max rel diff of 16x16 block means %.2e" % (
    n / 8 / st2.gpu_seconds / 1e6, f2.mean(dtype=np.float64), f1.mean(dtype=np.float64), (st1.extend_rays, st1.shadow_rays) == (st2.extend_rays, st2.shadow_rays),
    float(np.max(np.abs(f1[:528, :].reshape(33, 16, 60, 16, 3).mean(axis=(1, 3)) - f2[:528, :].reshape(33, 16, 60, 16, 3).mean(axis=(1, 3))) / (f2.mean() + 1e-9)))), flush=True)
fp, stp = g.render("ptdirect", n, W, H, max_num_vertices=6, seed=3)
print("c3 ptdirect m 6 Mpaths/s %.1f mean %.5f" % (n / stp.gpu_seconds / 1e6, fp.mean(dtype=np.float64)), flush=True)
