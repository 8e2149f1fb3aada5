"""Is the bdpt bench line's device-timed value depressed by the nvidia-smi clock sampler? (bdpt: ~14 000 launches and a host sync per
batch; pt / ptdirect: CUDA graphs.) Times the same renders with and without `nvidia-smi -lms 200` running beside them."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from nanogi_b200 import capi

def run(workload, spp, sampler):
    gen, renderer, W, H, _, m, desc = bench.WORKLOADS[workload]
    sd = bench.build_scene(workload, W / H)
    scene = capi.GpuScene(sd, 0)
    film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    n = W * H * spp
    for i in range(2):
        scene.render_device(film.data_ptr(), stream.cuda_stream, renderer, n, W, H, max_num_vertices=m, seed=1000 + i, film_norm_samples=n)
    torch.cuda.synchronize()
    out = []
    for label, on in (("without", False), ("with", True), ("without", False), ("with", True)):
        if on != sampler and sampler is not None:
            pass
        p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + bench.ClockSampler.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                             stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) if on else None
        time.sleep(0.5)
        t0 = time.perf_counter()
        for i in range(2):
            scene.render_device(film.data_ptr(), stream.cuda_stream, renderer, n, W, H, max_num_vertices=m, seed=2000 + i, film_norm_samples=n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if p:
            p.terminate(); p.wait()
        out.append((label, round(2 * n / dt / 1e6, 1)))
    scene.close()
    print(workload, spp, "spp, Mpaths/s", out, flush=True)

run("c2bdpt", 256, None)
run("c2", 256, None)
