#!/bin/bash
# A/B of alternative builds of the CUDA module (build/*.so) on the bdpt throughput table
OUT=gpurun_out; mkdir -p $OUT
for lib in build/*.so; do echo "== $lib"; NGI_GPU_LIB=$PWD/$lib timeout 300 python tools/bdpt_time.py 2>&1 | grep -E "m (3|6|-1) "; done | tee $OUT/s29_ab.txt
