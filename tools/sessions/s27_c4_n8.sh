#!/bin/bash
# 8 GPUs: BASELINE config[3] as named — C4 (10.5M triangles) ptdirect and pt, 1920x1080, 8 x 512 = 4096 spp, one process per GPU + one NCCL film reduce
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/s27_gpus.txt
timeout 300 python bench.py --gpus 8 --workload c4 --spp 512 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s27_bench_c4_n8.json 2> $OUT/s27_bench_c4_n8.err
timeout 300 python bench.py --gpus 8 --workload c4pt --spp 512 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s27_bench_c4pt_n8.json 2> $OUT/s27_bench_c4pt_n8.err
for f in $OUT/s27_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); print(sys.argv[1], j["n_gpus"], round(j["value"], 1), round(j["e2e"]["value"], 1), j["ms_per_step"], j.get("mrays_per_s"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -3 $OUT/s27_bench_c4_n8.err
