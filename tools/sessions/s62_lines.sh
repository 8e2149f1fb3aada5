#!/bin/bash
# round 2, session 62 (1 GPU): the default bench line (C3) and C2 with profiles/r02_ncu_metrics.json of THIS build in place (roofline.ncu / traffic filled)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py > $OUT/f5_bench_c3.json 2> $OUT/f5_bench_c3.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu > $OUT/f5_bench_c2.json 2> $OUT/f5_bench_c2.err
python - <<'PY'
import json
for f in ("f5_bench_c3", "f5_bench_c2"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "cpu", (j.get("cpu_baseline") or {}).get("value"), "frac", j["roofline"]["frac"], "traffic", j["roofline"].get("traffic"), "ncu", bool(j["roofline"].get("ncu")), j.get("film_mean"), (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "ERR", e)
PY
