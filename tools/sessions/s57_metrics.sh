#!/bin/bash
# round 2, session 57 (1 GPU): smoke(); ncu captures of the trace kernels on C2 and C4 for profiles/r02_ncu_metrics.json
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/s57_smoke.log 2>&1; tail -5 $OUT/s57_smoke.log
for wl in c2 c4; do
  spp=64; [ $wl = c4 ] && spp=32
  rm -f $OUT/s57_iter_$wl.txt
  NGI_LANES=1 NGI_ITER_LOG=$OUT/s57_iter_$wl.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
      -f -o $OUT/s57_prof_$wl python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload $wl --spp $spp --no-cpu > $OUT/s57_prof_$wl.log 2>&1
  ls -la $OUT/s57_prof_$wl.ncu-rep
done
