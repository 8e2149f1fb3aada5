#!/usr/bin/env python
"""GPU-box experiment: sweep the dynamic-fetch / postponing thresholds of the persistent trace kernels."""
import itertools
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, build_scene  # noqa: E402
from nanogi_b200 import capi  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
refills = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,4,8,16").split(",")]
trimins = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0,4,8,16").split(",")]
gen, renderer, W, H, _, m, desc = WORKLOADS[wl]
sd = build_scene(wl, W / H)
n = W * H * spp
print(desc, "spp", spp, flush=True)
for rf, tm in itertools.product(refills, trimins):
    os.environ["NGI_TRACE_REFILL_MIN"] = str(rf)
    os.environ["NGI_TRACE_TRI_MIN"] = str(tm)
    g = capi.GpuScene(sd, 0)
    g.render(renderer, n // 4, W, H, max_num_vertices=m, seed=1)
    f, st = g.render(renderer, n, W, H, max_num_vertices=m, seed=2)
    f2, s2 = g.render(renderer, n, W, H, max_num_vertices=m, seed=2, flags=capi.RENDER_TIME_KERNELS)
    print(f"refill_min={rf:2d} tri_min={tm:2d}  {n / st.gpu_seconds / 1e6:8.1f} Mpaths/s  {(st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e9:6.3f} Grays/s | "
          f"logic {s2.logic_kernel_seconds * 1e3:7.1f} ms extend {s2.extend_kernel_seconds * 1e3:7.1f} ms shadow {s2.shadow_kernel_seconds * 1e3:7.1f} ms | mean {f.mean():.5f}", flush=True)
    g.close()
