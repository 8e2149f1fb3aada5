#!/bin/bash
# Round-end check at HEAD (1 GPU): full GPU test suite, smoke(), both bench arms on the headline workload, C3, bdpt table + ncu of the bdpt stages
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/g_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/g_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/g_bench_ref_c2.json 2> $OUT/g_bench_ref_c2.err
timeout 600 python bench.py > $OUT/g_bench_c2.json 2> $OUT/g_bench_c2.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 3 --e2e-steps 1 > $OUT/g_bench_c3.json 2> $OUT/g_bench_c3.err
timeout 300 python tools/bdpt_time.py > $OUT/g_bdpt_wave.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bdw_contrib|k_bdw_shadow|k_bdw_start|k_bdw_count|k_bdw_expand|DeviceRadixSort' -s 8 -c 8 -f -o $OUT/prof_g_bdw1 python tools/bdpt_prof.py > $OUT/prof_g_bdw1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bdw_extend|k_bdw_step' -s 46 -c 4 -f -o $OUT/prof_g_bdw2 python tools/bdpt_prof.py > $OUT/prof_g_bdw2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/g_launches_bdw.csv python tools/bdpt_prof.py > $OUT/g_launches_bdw.log 2>&1
tail -4 $OUT/g_pytest.log; cat $OUT/g_smoke.log | tail -2; cat $OUT/g_bdpt_wave.txt
python - <<'PY'
import json
for f in ("g_bench_ref_c2", "g_bench_c2", "g_bench_c3"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), round(j.get("mrays_per_s") or 0, 1), round(j["e2e"]["value"], 2), (j.get("cpu_baseline") or {}).get("value"), (j.get("roofline") or {}).get("frac"), (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "ERR", e)
PY
