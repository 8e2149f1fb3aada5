#!/bin/bash
# GPU session 5: full GPU suite (incl. CLI vs the reference binary), both bench arms, image acceptance against nanogi's own
# code, the 10M-triangle C4 workload (bench + ncu).
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/s5_pytest.log 2>&1
( timeout 600 python -m pytest tests/test_reference_pin.py tests/test_golden.py -x -q -m "not gpu" ) > $OUT/s5_pytest_pin.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/s5_bench_ref_c2.json 2> $OUT/s5_bench_ref_c2.err
timeout 600 python bench.py > $OUT/s5_bench_c2.json 2> $OUT/s5_bench_c2.err
timeout 900 python tools/image_parity.py --scene cornell_box --renderer pt --size 128 --spp 64 -m 8 --seeds 32 > $OUT/s5_image_c1_pt.json 2> $OUT/s5_image_c1_pt.err
timeout 900 python tools/image_parity.py --scene cornell_spheres --renderer ptdirect --size 128 --spp 64 -m -1 --seeds 32 > $OUT/s5_image_c2_ptdirect.json 2> $OUT/s5_image_c2_ptdirect.err
timeout 900 python tools/image_parity.py --scene cornell_spheres --renderer ltdirect --size 128 --spp 64 -m -1 --seeds 16 > $OUT/s5_image_c2_ltdirect.json 2> $OUT/s5_image_c2_ltdirect.err
timeout 900 python bench.py --workload c4 --spp 64 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s5_bench_c4.json 2> $OUT/s5_bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 2 \
    -f -o $OUT/prof_s5c4 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c4 --spp 16 --no-cpu > $OUT/prof_s5c4.log 2>&1
tail -3 $OUT/s5_pytest.log; tail -2 $OUT/s5_pytest_pin.log
python - <<'PY'
import json
for f in ("s5_bench_ref_c2", "s5_bench_c2", "s5_bench_c4"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), j["e2e"]["value"], j.get("cpu_baseline"), (j.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
for f in ("s5_image_c1_pt", "s5_image_c2_ptdirect", "s5_image_c2_ltdirect"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, j["cpu_side"][:20], j["rel_rmse_clamped_gpu"], j["rel_rmse_clamped_oracle"], j["rel_rmse_clamped_diff_pct_of_oracle"], j["rel_rmse_clamped_diff_standard_error_pct"], j["block_z_max"], j["paired_replay"]["block_z_max"])
    except Exception as e:
        print(f, "ERR", e)
PY
