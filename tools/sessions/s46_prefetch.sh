#!/bin/bash
# round 2, session 46: software prefetch in the trace loop (next node -> L1, first triangle -> L1, ray records of a chunk -> L2) and
# 40 warps / SM (48 registers); GPU tests of the pieces added since s45 (bridge binary, loader fixes).
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag lib workload spp env...
  tag=$1; lib=$2; wl=$3; spp=$4; shift 4
  env "$@" NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s46_${wl}_${tag}.json 2> $OUT/s46_${wl}_${tag}.err
  python - $OUT/s46_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s | extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4),
          "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4), "| extend Grays/s", round(j["roofline"]["grays_per_s"], 3))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
P=$PWD
{
for v in base pf1 pf2 pf4 pf5 pf7 mb20 pf5_mb20; do run $v $P/build/$v.so c3 512 X=1; done
for v in base pf1 pf5 pf7; do run $v $P/build/$v.so c2 512 X=1; done
for v in base pf5 pf7; do run $v $P/build/$v.so c4 64 X=1; done
} | tee $OUT/s46_ab.txt
( time timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_gpu_branches.py tests/test_multi_gpu.py -m gpu -q ) > $OUT/s46_pytest.log 2>&1
tail -6 $OUT/s46_pytest.log
