#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bdpt" > $OUT/s26_pytest.log 2>&1; echo "pytest rc $?" >> $OUT/s26_pytest.log
tail -3 $OUT/s26_pytest.log
for K in 1 2 3 4; do echo "streams $K"; NGI_BDPT_STREAMS=$K timeout 300 python tools/bdpt_time.py 2>&1 | grep -E "m (6|-1) "; done | tee $OUT/s26_bdpt_streams.txt
