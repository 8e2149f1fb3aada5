#!/bin/bash
# round 2, session 56: run-time knobs on the final trace kernels — lanes (concurrent wavefront pipelines), wave capacity
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag workload spp extra-args env...
  tag=$1; wl=$2; spp=$3; extra=$4; shift 4
  env "$@" timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu $extra > $OUT/s56_${wl}_${tag}.json 2> $OUT/s56_${wl}_${tag}.err
  python - $OUT/s56_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s ms/step", round(j["ms_per_step"], 1), "iters", j["wave_iterations"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
{
run l2 c3 1024 "" X=1
run l1 c3 1024 "" NGI_LANES=1
run l3 c3 1024 "" NGI_LANES=3
run l4 c3 1024 "" NGI_LANES=4
run l3_w4m c3 1024 "--wave-capacity 4194304" NGI_LANES=3
run l2_w16m c3 1024 "--wave-capacity 16777216" X=1
run l2 c2 1024 "" X=1
run l3 c2 1024 "" NGI_LANES=3
run l2 c4 128 "" X=1
run l3 c4 128 "" NGI_LANES=3
} | tee $OUT/s56_ab.txt
