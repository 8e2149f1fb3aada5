#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1; do echo "NGI_TRACE_SPREAD=$v"; NGI_TRACE_SPREAD=$v timeout 300 python tools/bdpt_time.py 2>&1 | grep -E "m (3|6|-1) "; NGI_TRACE_SPREAD=$v NGI_BDPT_BATCH=524288 timeout 300 python tools/bdpt_time.py 2>&1 | grep -E "m (-1) " | sed 's/^/batch 2^19: /'; done | tee $OUT/s35_spread_bdpt.txt
bash tools/sessions/s31_env.sh NGI_TRACE_SPREAD "0 1" "c2 c1"
