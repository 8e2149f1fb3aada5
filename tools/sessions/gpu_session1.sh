#!/bin/bash
# One GPU-box session: tests, benches (both arms), image acceptance, launch list + full ncu capture.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/s1_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/s1_pytest.log 2>&1
timeout 600 python bench.py > $OUT/s1_bench_c2.json 2> $OUT/s1_bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/s1_bench_ref_c2.json 2> $OUT/s1_bench_ref_c2.err
timeout 600 python bench.py --workload c3 --steps 2 --warmup 3 --e2e-steps 1 > $OUT/s1_bench_c3.json 2> $OUT/s1_bench_c3.err
timeout 600 python tools/image_parity.py --scene cornell_box --renderer pt --size 128 --spp 64 -m 8 > $OUT/s1_image_c1_pt.json 2> $OUT/s1_image_c1_pt.err
timeout 600 python tools/image_parity.py --scene cornell_spheres --renderer ptdirect --size 128 --spp 64 -m -1 > $OUT/s1_image_c2_ptdirect.json 2> $OUT/s1_image_c2_ptdirect.err
bash tools/gpu_profile.sh s1c2 c2 64 60
tail -3 $OUT/s1_pytest.log; cat $OUT/s1_bench_c2.json | cut -c1-600
