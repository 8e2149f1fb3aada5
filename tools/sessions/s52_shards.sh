#!/bin/bash
# round 2, session 52 (1 GPU): shard emulation on one GPU (film of 8 / 4 shards summed = film of the whole job?) + smem carve-out A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/check_shards.py --workload c3 --shards 8 > $OUT/s52_shards_c3.json 2> $OUT/s52_shards_c3.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/s52_shards_c3.json").read().strip().splitlines()[-1])
for k, v in j.items():
    print(k, v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "shard_means"})
print(j["shards_8"]["shard_means"])
PY
run() {  # tag workload spp env...
  tag=$1; wl=$2; spp=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s52_${wl}_${tag}.json 2> $OUT/s52_${wl}_${tag}.err
  python - $OUT/s52_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s | extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4), "film mean", j["film_mean"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
{
run dflt c3 512 X=1
run co8  c3 512 NGI_TRACE_CARVEOUT=8
run co16 c3 512 NGI_TRACE_CARVEOUT=16
run co25 c3 512 NGI_TRACE_CARVEOUT=25
run co50 c3 512 NGI_TRACE_CARVEOUT=50
} | tee $OUT/s52_ab.txt
