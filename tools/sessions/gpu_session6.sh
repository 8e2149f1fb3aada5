#!/bin/bash
# GPU session 6: sector-record path state (A/B against session 5), logic-kernel ncu, C3 image acceptance.
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/s6_pytest.log 2>&1
timeout 600 python bench.py --no-cpu > $OUT/s6_bench_c2.json 2> $OUT/s6_bench_c2.err
timeout 600 python bench.py --no-cpu --workload c3 --steps 2 --warmup 3 --e2e-steps 1 > $OUT/s6_bench_c3.json 2> $OUT/s6_bench_c3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_classify|k_surface|k_eye' -s 60 -c 5 \
    -f -o $OUT/prof_s6c2 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c2 --spp 64 --no-cpu > $OUT/prof_s6c2.log 2>&1
timeout 1200 python tools/image_parity.py --scene instanced_spheres --renderer ptdirect --width 256 --height 144 --spp 64 -m -1 --seeds 8 --cpu-side oracle --oracle-ref-spp 1024 --paired-spp 128 > $OUT/s6_image_c3_ptdirect.json 2> $OUT/s6_image_c3_ptdirect.err
tail -3 $OUT/s6_pytest.log
python - <<'PY'
import json
for f in ("s6_bench_c2", "s6_bench_c3"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, round(j["value"], 2), j["e2e"]["value"], {k: round(v["avg_launch_ms"], 4) for k, v in j["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
try:
    j = json.loads(open("gpurun_out/s6_image_c3_ptdirect.json").read().strip().splitlines()[-1])
    print("c3 image", j["rel_rmse_clamped_gpu"], j["rel_rmse_clamped_oracle"], j["rel_rmse_clamped_diff_pct_of_oracle"], j["rel_rmse_clamped_diff_standard_error_pct"], j["block_z_max"], j["block_z_frac_gt3"], j["paired_replay"]["block_z_max"], j["paired_replay"]["block_z_frac_gt3"])
except Exception as e:
    print("c3 image ERR", e)
PY
