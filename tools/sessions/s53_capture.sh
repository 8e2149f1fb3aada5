#!/bin/bash
# round 2, session 53 (1 GPU): HEAD after the stream-ordering fix — GPU suite, default bench line + reference arm, launch list + ncu
# capture of the same command (-> profiles/r02_ncu_metrics.json)
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/f3_pytest.log 2>&1
tail -4 $OUT/f3_pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/f3_bench_ref_c3.json 2> $OUT/f3_bench_ref_c3.err
rm -f $OUT/f3_iter.txt
NGI_LANES=1 NGI_ITER_LOG=$OUT/f3_iter.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
    -f -o $OUT/f3_prof_c3 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 64 --no-cpu > $OUT/f3_prof_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/f3_launches_c3.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 8 --no-cpu > $OUT/f3_launches_c3.log 2>&1
timeout 900 python bench.py > $OUT/f3_bench_c3.json 2> $OUT/f3_bench_c3.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu > $OUT/f3_bench_c2.json 2> $OUT/f3_bench_c2.err
python - <<'PY'
import json
for f in ("f3_bench_ref_c3", "f3_bench_c3", "f3_bench_c2"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), round(j.get("mrays_per_s") or 0, 1), "e2e", round(j["e2e"]["value"], 2), "cpu", (j.get("cpu_baseline") or {}).get("value"), (j.get("cpu_baseline") or {}).get("sample"), "frac", (j.get("roofline") or {}).get("frac"), "film", j.get("film_mean"), (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "ERR", e)
PY
