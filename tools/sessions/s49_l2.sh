#!/bin/bash
# round 2, session 49: L2 persistence window over BVH nodes + triangles (NGI_L2_PERSIST_MB) and the pooled fp64 film accumulator (e2e)
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag lib workload spp env...
  tag=$1; lib=$2; wl=$3; spp=$4; shift 4
  env "$@" NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 2 --no-cpu > $OUT/s49_${wl}_${tag}.json 2> $OUT/s49_${wl}_${tag}.err
  python - $OUT/s49_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s e2e", round(j["e2e"]["value"], 1), "| extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4),
          "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4), "| e2e parts", {a: round(b, 4) for a, b in j["e2e"]["breakdown_s_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
  grep "L2 persist" $OUT/s49_${wl}_${tag}.err | head -1
}
P=$PWD
{
run head  $P/nanogi_b200/libnanogi_gpu.so c3 512 X=1
run l2_0  $P/build/l2.so c3 512 NGI_L2_PERSIST_MB=0
run l2_32 $P/build/l2.so c3 512 NGI_L2_PERSIST_MB=32 NGI_L2_PERSIST_VERBOSE=1
run l2_64 $P/build/l2.so c3 512 NGI_L2_PERSIST_MB=64 NGI_L2_PERSIST_VERBOSE=1
run l2_96 $P/build/l2.so c3 512 NGI_L2_PERSIST_MB=96 NGI_L2_PERSIST_VERBOSE=1
run head  $P/nanogi_b200/libnanogi_gpu.so c2 512 X=1
run l2_0  $P/build/l2.so c2 512 NGI_L2_PERSIST_MB=0
run l2_64 $P/build/l2.so c2 512 NGI_L2_PERSIST_MB=64
run l2_0  $P/build/l2.so c4 64 NGI_L2_PERSIST_MB=0
run l2_64 $P/build/l2.so c4 64 NGI_L2_PERSIST_MB=64 NGI_L2_PERSIST_VERBOSE=1
run l2_96 $P/build/l2.so c4 64 NGI_L2_PERSIST_MB=96
run l2_0  $P/build/l2.so c1 64 NGI_L2_PERSIST_MB=0
} | tee $OUT/s49_ab.txt
