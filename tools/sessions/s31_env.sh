#!/bin/bash
# A/B of an environment knob on the C3 / C2 bench lines: usage s31_env.sh VAR "v1 v2 ..." "workloads"
OUT=gpurun_out; mkdir -p $OUT; VAR=$1; VALS=$2; WLS=${3:-"c3 c2"}
for wl in $WLS; do for v in $VALS; do
  spp=""; [ $wl = c4 ] && spp="--spp 128"
  if [ "$v" = "unset" ]; then envs=""; else envs="$VAR=$v"; fi
  env $envs timeout 400 python bench.py --workload $wl $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s31_${wl}_${VAR}_${v}.json 2> $OUT/s31_${wl}_${VAR}_${v}.err
  python - $OUT/s31_${wl}_${VAR}_${v}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4), "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done; done | tee $OUT/s31_${VAR}.txt
