"""bdpt, 256 spp on C2's scene: 12 consecutive renders on ONE scene handle, then 3 on a fresh one — wall time, the module's own
device time (CUDA events around the render) and launch count per render."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from nanogi_b200 import capi

gen, renderer, W, H, _, m, desc = bench.WORKLOADS["c2bdpt"]
sd = bench.build_scene("c2bdpt", W / H)
film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream()
n = W * H * 256
for tag, reps in (("handle A", 12), ("handle B (fresh)", 3)):
    scene = capi.GpuScene(sd, 0)
    for i in range(reps):
        free0, _ = torch.cuda.mem_get_info()
        t0 = time.perf_counter()
        st = scene.render_device(film.data_ptr(), stream.cuda_stream, renderer, n, W, H, max_num_vertices=m, seed=2000 + (i & 1), film_norm_samples=n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{tag} render {i}: wall {dt:.3f} s = {n / dt / 1e6:.1f} Mpaths/s, device {st.gpu_seconds:.3f} s, launches {st.kernel_launches}, batches {st.wave_iterations}, "
              f"free before {free0 / 2**30:.1f} GiB", flush=True)
    scene.close()
