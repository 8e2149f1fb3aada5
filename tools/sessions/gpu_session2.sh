#!/bin/bash
# GPU session 2: pooled allocation A/B on the e2e path, image acceptance (robust + paired), C3 source-level profile.
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/s2_pytest.log 2>&1
NGI_POOL_ALLOC=0 timeout 300 python bench.py --no-cpu > $OUT/s2_bench_c2_nopool.json 2> $OUT/s2_bench_c2_nopool.err
NGI_POOL_ALLOC=1 timeout 300 python bench.py --no-cpu > $OUT/s2_bench_c2_pool.json 2> $OUT/s2_bench_c2_pool.err
NGI_POOL_ALLOC=0 timeout 300 python bench.py --no-cpu --workload c3 --steps 1 --warmup 1 --e2e-steps 2 > $OUT/s2_bench_c3_nopool.json 2> $OUT/s2_bench_c3_nopool.err
NGI_POOL_ALLOC=1 timeout 300 python bench.py --no-cpu --workload c3 --steps 1 --warmup 1 --e2e-steps 2 > $OUT/s2_bench_c3_pool.json 2> $OUT/s2_bench_c3_pool.err
timeout 900 python tools/image_parity.py --scene cornell_box --renderer pt --size 128 --spp 64 -m 8 --seeds 32 > $OUT/s2_image_c1_pt.json 2> $OUT/s2_image_c1_pt.err
timeout 900 python tools/image_parity.py --scene cornell_spheres --renderer ptdirect --size 128 --spp 64 -m -1 --seeds 32 > $OUT/s2_image_c2_ptdirect.json 2> $OUT/s2_image_c2_ptdirect.err
timeout 900 python tools/image_parity.py --scene cornell_spheres --renderer pt --size 128 --spp 64 -m -1 --seeds 32 > $OUT/s2_image_c2_pt.json 2> $OUT/s2_image_c2_pt.err
# C3: full capture with source of the trace kernels in steady state
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_surface' -s 90 -c 3 \
    -f -o $OUT/prof_s2c3 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 32 --no-cpu > $OUT/prof_s2c3.log 2>&1
tail -2 $OUT/s2_pytest.log
for f in $OUT/s2_bench_*.json; do python - "$f" <<'PY'
import json, sys
j = json.load(open(sys.argv[1])); print(sys.argv[1], round(j["value"], 1), j["e2e"]["value"], j["e2e"].get("breakdown_s_per_step"))
PY
done
