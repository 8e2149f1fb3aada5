#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --workload c2bdpt > $OUT/s36_bench_c2bdpt.json 2> $OUT/s36_bench_c2bdpt.err
timeout 600 python bench.py --workload c2bdpt --impl reference --steps 3 --warmup 1 > $OUT/s36_bench_ref_c2bdpt.json 2> $OUT/s36_bench_ref_c2bdpt.err
tail -2 $OUT/s36_bench_c2bdpt.err; cat $OUT/s36_bench_c2bdpt.json | cut -c1-3000; cat $OUT/s36_bench_ref_c2bdpt.json | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bdpt" 2>&1 | tail -2
