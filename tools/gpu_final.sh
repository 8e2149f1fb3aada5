#!/bin/bash
# Round-end measurement session at HEAD (1 GPU): tests, both bench arms, every workload, launch list + full ncu capture, C5 on C4.
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/f_pytest.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/f_bench_ref_c2.json 2> $OUT/f_bench_ref_c2.err
timeout 600 python bench.py > $OUT/f_bench_c2.json 2> $OUT/f_bench_c2.err
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > $OUT/f_bench_c1.json 2> $OUT/f_bench_c1.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 3 --e2e-steps 1 > $OUT/f_bench_c3.json 2> $OUT/f_bench_c3.err
timeout 900 python bench.py --workload c4 --spp 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/f_bench_c4.json 2> $OUT/f_bench_c4.err
timeout 900 python bench.py --workload c4pt --spp 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/f_bench_c4pt.json 2> $OUT/f_bench_c4pt.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/f_launches_c2.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c2 --spp 8 --no-cpu > $OUT/f_launches_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_classify|k_surface|k_eye' -s 60 -c 5 \
    -f -o $OUT/f_prof_c2 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c2 --spp 64 --no-cpu > $OUT/f_prof_c2.log 2>&1
timeout 1200 python tools/raybench.py --scene c4 --rays 16777216 --check 16777216 > $OUT/f_raybench_c4_16M.json 2> $OUT/f_raybench_c4.err
tail -4 $OUT/f_pytest.log
python - <<'PY'
import json
for f in ("f_bench_ref_c2", "f_bench_c2", "f_bench_c1", "f_bench_c3", "f_bench_c4", "f_bench_c4pt"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), round(j.get("mrays_per_s") or 0, 1), round(j["e2e"]["value"], 2), (j.get("cpu_baseline") or {}).get("value"), (j.get("roofline") or {}).get("frac"), (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "ERR", e)
try:
    j = json.loads(open("gpurun_out/f_raybench_c4_16M.json").read().strip().splitlines()[-1])
    print({k: (round(v["grays_per_s"], 2), v["checked"], v["mismatches"]) for k, v in j["batches"].items()})
except Exception as e:
    print("raybench ERR", e)
PY
