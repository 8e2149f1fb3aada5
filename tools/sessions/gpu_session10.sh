#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/s10_pytest.log 2>&1
timeout 600 python - > $OUT/s10_bdpt_throughput.txt 2>&1 <<'PY'
import sys, os; sys.path.insert(0, ".")
import numpy as np
from nanogi_b200 import capi, scenes
from oracle import pyref, pyoracle
import time
for name, spec, w, h in (("c2", scenes.cornell_spheres(), 1024, 1024), ("c3", scenes.instanced_spheres(), 1920, 1080)):
    sd = scenes.to_scene_data(spec, w / h)
    g = capi.GpuScene(sd, 0)
    for r in ("ptdirect", "bdpt"):
        g.render(r, 1 << 22, w, h, seed=1)
        n = 1 << 26
        f, st = g.render(r, n, w, h, seed=2)
        print(name, r, "GPU Mpaths/s %.1f Mrays/s %.1f mean %.5g" % (n / st.gpu_seconds / 1e6, (st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6, f.mean()), flush=True)
    if name == "c2":
        orc = pyoracle.OracleScene(sd)
        t0 = time.time(); fo, so = orc.render("bdpt", 1 << 22, w, h, seed=3); dt = time.time() - t0
        print(name, "bdpt oracle port (%d cores) Mpaths/s %.2f mean %.5g" % (os.cpu_count(), (1 << 22) / dt / 1e6, fo.mean()), flush=True)
        if pyref.available():
            ref = pyref.RefScene(spec, w / h)
            t0 = time.time(); fr = ref.render("bdpt", 1 << 22, w, h, seed=3, num_threads=os.cpu_count()); dt = time.time() - t0
            print(name, "bdpt nanogi's own code (%d cores) Mpaths/s %.2f mean %.5g" % (os.cpu_count(), (1 << 22) / dt / 1e6, fr.mean()), flush=True)
    g.close()
PY
timeout 900 python tools/image_parity.py --scene cornell_spheres --renderer bdpt --size 128 --spp 64 -m -1 --seeds 16 > $OUT/s10_image_c2_bdpt.json 2> $OUT/s10_image_c2_bdpt.err
tail -4 $OUT/s10_pytest.log; cat $OUT/s10_bdpt_throughput.txt
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/s10_image_c2_bdpt.json").read().strip().splitlines()[-1])
    print("bdpt image", j["cpu_side"][:20], j["rel_rmse_clamped_gpu"], j["rel_rmse_clamped_oracle"], j["rel_rmse_clamped_diff_pct_of_oracle"], j["rel_rmse_clamped_diff_standard_error_pct"], j["block_z_max"], j["block_z_frac_gt3"], j["paired_replay"]["block_z_max"], j["paired_replay"]["block_z_frac_gt3"])
except Exception as e:
    print("bdpt image ERR", e)
PY
