#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -x -q -k "bdpt" > $OUT/s34_pytest.log 2>&1; echo "pytest rc $?" >> $OUT/s34_pytest.log
tail -3 $OUT/s34_pytest.log
timeout 300 python tools/bdpt_time.py 2>&1 | grep -E "m (6|-1) "
timeout 300 compute-sanitizer --print-limit 5 python tools/bdw_small.py 2>&1 | tail -4
