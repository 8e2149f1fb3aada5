#!/bin/bash
# A/B of alternative builds of the CUDA module (build/*.so) on the C3 / C2 / C4 bench lines
OUT=gpurun_out; mkdir -p $OUT
for wl in c3 c4; do
for lib in build/*.so; do
  tag=$(basename $lib .so)
  spp=""; [ $wl = c4 ] && spp="--spp 128"
  NGI_GPU_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s30_${wl}_${tag}.json 2> $OUT/s30_${wl}_${tag}.err
  python - $OUT/s30_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done; done | tee $OUT/s30_ab.txt
