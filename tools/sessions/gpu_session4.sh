#!/bin/bash
# GPU session 4 (8 GPUs): weak scaling of the C3 and C2 workloads, one process per GPU + one NCCL film reduce.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/s4_gpus.txt
timeout 600 python bench.py --gpus 8 --workload c3 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s4_bench_c3_n8.json 2> $OUT/s4_bench_c3_n8.err
timeout 600 python bench.py --gpus 4 --workload c3 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s4_bench_c3_n4.json 2> $OUT/s4_bench_c3_n4.err
timeout 600 python bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > $OUT/s4_bench_c2_n8.json 2> $OUT/s4_bench_c2_n8.err
for f in $OUT/s4_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); print(sys.argv[1], j["n_gpus"], round(j["value"], 1), round(j["e2e"]["value"], 1), j["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -3 $OUT/s4_bench_c3_n8.err
