#!/bin/bash
# GPU box: wavefront bdpt after a change — parity tests, throughput table, ncu of the contribution stage
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-s23}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bdpt" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> $OUT/${TAG}_pytest.log
timeout 300 python tools/bdpt_time.py > $OUT/${TAG}_bdpt_wave.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bdw_contrib|k_bdw_pre' -s 1 -c 2 -f -o $OUT/prof_${TAG} python tools/bdpt_prof.py > $OUT/prof_${TAG}.log 2>&1
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_bdpt_wave.txt
