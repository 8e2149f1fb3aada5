#!/bin/bash
# round 2, session 54 (gpurun --gpus 8): after the stream-ordering fix — the product's reduce against patterns and torch.distributed.reduce
# on 8 ranks, then the C3 strong-scaling line at N = 8 (film mean must equal the 1-GPU film's: same sample set)
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/check_reduce.py > $OUT/n8v_reduce.json 2> $OUT/n8v_reduce.err
tail -1 $OUT/n8v_reduce.json
NCCL_DEBUG=VERSION timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/n8v_bench_c3_n8.json 2> $OUT/n8v_bench_c3_n8.err
NCCL_DEBUG=VERSION timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 \
    bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/n8v_bench_c3_n2.json 2> $OUT/n8v_bench_c3_n2.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > $OUT/n8v_bench_c3_n1.json 2> $OUT/n8v_bench_c3_n1.err
python - <<'PY'
import json
v = {}
for n in (1, 2, 8):
    try:
        j = json.loads([l for l in open(f"gpurun_out/n8v_bench_c3_n{n}.json").read().splitlines() if l.startswith("{")][-1])
        v[n] = j
        print("N =", n, round(j["value"], 1), "Mpaths/s", j["scaling"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1), "film mean", j["film_mean"], "clocks", (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print("N =", n, "ERR", e)
for n in (2, 8):
    if 1 in v and n in v:
        print("strong-scaling efficiency at N = %d: %.3f (device-timed), %.3f (e2e)" % (n, v[n]["value"] / (n * v[1]["value"]), v[n]["e2e"]["value"] / (n * v[1]["e2e"]["value"])))
PY
