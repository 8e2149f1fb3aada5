#!/bin/bash
# round 2, session 63 (gpurun --gpus 8): the C3 strong-scaling line at N = 8 on the final build (film mean must equal the 1-GPU film's)
OUT=gpurun_out; mkdir -p $OUT
NCCL_DEBUG=VERSION timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/f5_bench_c3_n8.json 2> $OUT/f5_bench_c3_n8.err
python - <<'PY'
import json
try:
    j = json.loads([l for l in open("gpurun_out/f5_bench_c3_n8.json").read().splitlines() if l.startswith("{")][-1])
    print("N = 8", round(j["value"], 1), "Mpaths/s", j["scaling"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1), "film mean", j["film_mean"], "clocks", (j.get("clocks") or {}).get("reasons"))
except Exception as e:
    print("ERR", e)
PY
tail -3 $OUT/f5_bench_c3_n8.err
