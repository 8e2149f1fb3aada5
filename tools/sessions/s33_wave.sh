#!/bin/bash
# wave capacity beyond 4 Mi slots per lane (C2 / C3 bench lines)
OUT=gpurun_out; mkdir -p $OUT
for wl in c2 c3; do for P in 4194304 8388608 16777216; do
  timeout 400 python bench.py --workload $wl --wave-capacity $P --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s33_${wl}_${P}.json 2> $OUT/s33_${wl}_${P}.err
  python - $OUT/s33_${wl}_${P}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4), "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done; done | tee $OUT/s33_wave.txt
