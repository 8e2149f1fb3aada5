#!/bin/bash
# round 2, session 51: 96-byte nodes / 64-byte triangles fetched with 256-bit loads (3 + 2 load instructions instead of 5 + 3)
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag lib workload spp env...
  tag=$1; lib=$2; wl=$3; spp=$4; shift 4
  env "$@" NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s51_${wl}_${tag}.json 2> $OUT/s51_${wl}_${tag}.err
  python - $OUT/s51_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s | extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4),
          "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4), "| scene MB", round(j["config"]["scene_device_bytes"] / 1e6, 1))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
P=$PWD
{
run head  $P/nanogi_b200/libnanogi_gpu.so c3 512 X=1
run ld256 $P/build/ld256.so c3 512 X=1
run head  $P/nanogi_b200/libnanogi_gpu.so c2 512 X=1
run ld256 $P/build/ld256.so c2 512 X=1
run head  $P/nanogi_b200/libnanogi_gpu.so c4 64 X=1
run ld256 $P/build/ld256.so c4 64 X=1
} | tee $OUT/s51_ab.txt
NGI_GPU_LIB=$P/build/ld256.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/s51_pytest.log 2>&1
tail -3 $OUT/s51_pytest.log
