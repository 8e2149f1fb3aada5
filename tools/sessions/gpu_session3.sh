#!/bin/bash
# GPU session 3: full GPU test suite incl. lt / ltdirect / E.area / CLI, hot-path regression bench, light-tracing throughput.
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/s3_pytest.log 2>&1
timeout 300 python bench.py --no-cpu > $OUT/s3_bench_c2.json 2> $OUT/s3_bench_c2.err
timeout 300 python - > $OUT/s3_lt_throughput.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
from nanogi_b200 import capi, scenes
sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
g = capi.GpuScene(sd, 0)
for r in ("pt", "ptdirect", "lt", "ltdirect"):
    g.render(r, 1 << 24, 1024, 1024, seed=1)
    f, st = g.render(r, 1 << 28, 1024, 1024, seed=2)
    print(r, "Mpaths/s %.1f Mrays/s %.1f mean %.5g" % ((1 << 28) / st.gpu_seconds / 1e6, (st.extend_rays + st.shadow_rays) / st.gpu_seconds / 1e6, f.mean()))
PY
tail -5 $OUT/s3_pytest.log; cat $OUT/s3_lt_throughput.txt; python - <<'PY'
import json
j = json.load(open("gpurun_out/s3_bench_c2.json")); print(j["value"], j["e2e"]["value"], j["roofline"]["frac"])
PY
