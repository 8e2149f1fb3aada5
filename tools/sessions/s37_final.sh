#!/bin/bash
# Last check of the round at HEAD (1 GPU): full GPU suite, smoke(), default bench (both arms), bdpt bench line
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/h_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/h_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/h_bench_ref_c2.json 2> $OUT/h_bench_ref_c2.err
timeout 600 python bench.py > $OUT/h_bench_c2.json 2> $OUT/h_bench_c2.err
timeout 600 python bench.py --workload c2bdpt --no-cpu > $OUT/h_bench_c2bdpt.json 2> $OUT/h_bench_c2bdpt.err
tail -4 $OUT/h_pytest.log; tail -1 $OUT/h_smoke.log
python - <<'PY'
import json
for f in ("h_bench_ref_c2", "h_bench_c2", "h_bench_c2bdpt"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), round(j.get("mrays_per_s") or 0, 1), round(j["e2e"]["value"], 2), (j.get("cpu_baseline") or {}).get("value"), (j.get("roofline") or {}).get("frac"), (j.get("clocks") or {}).get("reasons"), j.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
