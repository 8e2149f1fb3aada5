#!/bin/bash
# round 2, final 1-GPU session at HEAD (after the SFU shading arithmetic): GPU suite, smoke(), both bench arms on the default workload
# (C3), the other BASELINE configs, launch list + full ncu captures of the trace kernels on C3 / C2 / C4 (-> profiles/r02_ncu_metrics.json
# through tools/ncu_metrics.py), C5 on C3.
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/f4_pytest.log 2>&1
tail -4 $OUT/f4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f4_smoke.log 2>&1; tail -3 $OUT/f4_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/f4_bench_ref_c3.json 2> $OUT/f4_bench_ref_c3.err
for wl in c3 c2 c4; do
  spp=64; [ $wl = c4 ] && spp=32
  rm -f $OUT/f4_iter_$wl.txt
  NGI_LANES=1 NGI_ITER_LOG=$OUT/f4_iter_$wl.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
      -f -o $OUT/f4_prof_$wl python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload $wl --spp $spp --no-cpu > $OUT/f4_prof_$wl.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/f4_launches_c3.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 8 --no-cpu > $OUT/f4_launches_c3.log 2>&1
timeout 900 python bench.py > $OUT/f4_bench_c3.json 2> $OUT/f4_bench_c3.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu > $OUT/f4_bench_c2.json 2> $OUT/f4_bench_c2.err
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > $OUT/f4_bench_c1.json 2> $OUT/f4_bench_c1.err
timeout 900 python bench.py --workload c4 --spp 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/f4_bench_c4.json 2> $OUT/f4_bench_c4.err
timeout 900 python bench.py --workload c4pt --spp 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/f4_bench_c4pt.json 2> $OUT/f4_bench_c4pt.err
timeout 600 python bench.py --workload c2bdpt --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/f4_bench_c2bdpt.json 2> $OUT/f4_bench_c2bdpt.err
timeout 900 python tools/raybench.py --scene c3 --rays 16777216 --check 4194304 > $OUT/f4_raybench_c3_16M.json 2> $OUT/f4_raybench_c3.err
python - <<'PY'
import json
for f in ("f4_bench_ref_c3", "f4_bench_c3", "f4_bench_c2", "f4_bench_c1", "f4_bench_c4", "f4_bench_c4pt", "f4_bench_c2bdpt"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), round(j.get("mrays_per_s") or 0, 1), "e2e", round(j["e2e"]["value"], 2), "cpu", (j.get("cpu_baseline") or {}).get("value"), "frac", (j.get("roofline") or {}).get("frac"), "film", j.get("film_mean"), (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "ERR", e)
try:
    j = json.loads(open("gpurun_out/f4_raybench_c3_16M.json").read().strip().splitlines()[-1])
    print("raybench c3", {k: (round(v["grays_per_s"], 2), v["checked"], v["mismatches"]) for k, v in j["batches"].items()})
except Exception as e:
    print("raybench ERR", e)
PY
