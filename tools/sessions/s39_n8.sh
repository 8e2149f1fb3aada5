#!/bin/bash
# 8 GPUs at HEAD: C3 weak scaling and the bdpt workload
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --gpus 8 --workload c3 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s39_bench_c3_n8.json 2> $OUT/s39_bench_c3_n8.err
timeout 300 python bench.py --gpus 8 --workload c2bdpt --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s39_bench_c2bdpt_n8.json 2> $OUT/s39_bench_c2bdpt_n8.err
for f in $OUT/s39_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); print(sys.argv[1], j["n_gpus"], round(j["value"], 1), round(j["e2e"]["value"], 1), j["ms_per_step"], j.get("mrays_per_s"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
