import sys; sys.path.insert(0, ".")
from nanogi_b200 import capi, scenes
sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
g = capi.GpuScene(sd, 0)
f, st = g.render("bdpt", 20000, 64, 64, seed=1, max_num_vertices=int(sys.argv[1]) if len(sys.argv) > 1 else -1)
print(f.mean(), st.extend_rays, st.shadow_rays, st.kernel_launches)
