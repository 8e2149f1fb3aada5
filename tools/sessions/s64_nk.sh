#!/bin/bash
# round 2, session 64 (gpurun --gpus N, N = $1): the C3 strong-scaling line at N GPUs on the final build
N=$1; OUT=gpurun_out; mkdir -p $OUT
NCCL_DEBUG=VERSION timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $N --steps 3 --warmup 3 > $OUT/f5_bench_c3_n$N.json 2> $OUT/f5_bench_c3_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads([l for l in open(f"gpurun_out/f5_bench_c3_n{n}.json").read().splitlines() if l.startswith("{")][-1])
    print("N =", n, round(j["value"], 1), "Mpaths/s", j["scaling"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1), "film mean", j["film_mean"], "clocks", (j.get("clocks") or {}).get("reasons"))
except Exception as e:
    print("ERR", e)
PY
