#!/bin/bash
# GPU box: first run of the wavefront bdpt — parity tests, throughput table (both forms), launch list
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/s20_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bdpt" > $OUT/s20_pytest.log 2>&1; echo "pytest rc $?" >> $OUT/s20_pytest.log
timeout 300 python tools/bdpt_time.py > $OUT/s20_bdpt_wave.txt 2>&1
timeout 300 python tools/bdpt_time.py --per-thread > $OUT/s20_bdpt_thread.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/s20_launches_bdw.csv python tools/bdpt_prof.py > $OUT/s20_launches_bdw.log 2>&1
tail -5 $OUT/s20_pytest.log; cat $OUT/s20_bdpt_wave.txt $OUT/s20_bdpt_thread.txt
