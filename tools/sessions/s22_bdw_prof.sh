#!/bin/bash
# GPU box: ncu --set full of the wavefront bdpt stages (second batch of the render)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bdw_contrib|k_bdw_shadow|k_bdw_start' -s 3 -c 3 -f -o $OUT/prof_bdw1 python tools/bdpt_prof.py > $OUT/prof_bdw1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bdw_extend|k_bdw_step' -s 46 -c 4 -f -o $OUT/prof_bdw2 python tools/bdpt_prof.py > $OUT/prof_bdw2.log 2>&1
ls -la $OUT/prof_bdw*
