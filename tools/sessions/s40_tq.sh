#!/bin/bash
# round 2, session 40: the shared-memory triangle backlog (ngi_trace_warp_tq) against the postponing loop of round 1.
# GPU tests with the new default, then A/B builds x tri_min on C3 (and C2), then one full ncu capture of the trace kernels on C3.
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/s40_pytest.log 2>&1
tail -3 $OUT/s40_pytest.log
run() {  # tag lib env...
  tag=$1; lib=$2; wl=$3; shift 3
  env "$@" NGI_GPU_LIB=$lib timeout 300 python bench.py --workload $wl --spp 512 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s40_${wl}_${tag}.json 2> $OUT/s40_${wl}_${tag}.err
  python - $OUT/s40_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s | extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4),
          "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4), "| extend Grays/s", round(j["roofline"]["grays_per_s"], 3))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
P=$PWD
{
run old      $P/build/old.so c3 X=1
run tq4_t8   $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=8
run tq4_t12  $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=12
run tq4_t16  $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=16
run tq4_t20  $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=20
run tq4_t24  $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=24
run tq2_t16  $P/build/tq2.so c3 NGI_TRACE_TRI_MIN=16
run tq6_t16  $P/build/tq6.so c3 NGI_TRACE_TRI_MIN=16
run tq6_t24  $P/build/tq6.so c3 NGI_TRACE_TRI_MIN=24
run r64_t16  $P/build/tq4r64.so c3 NGI_TRACE_TRI_MIN=16
run tq4_t16_r2 $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=16 NGI_TRACE_REFILL_MIN=2
run tq4_t16_r8 $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=16 NGI_TRACE_REFILL_MIN=8
run old      $P/build/old.so c2 X=1
run tq4_t8   $P/nanogi_b200/libnanogi_gpu.so c2 NGI_TRACE_TRI_MIN=8
run tq4_t16  $P/nanogi_b200/libnanogi_gpu.so c2 NGI_TRACE_TRI_MIN=16
run tq4_t24  $P/nanogi_b200/libnanogi_gpu.so c2 NGI_TRACE_TRI_MIN=24
} | tee $OUT/s40_ab.txt
NGI_TRACE_TRI_MIN=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
    -f -o $OUT/s40_prof_c3 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 64 --no-cpu > $OUT/s40_prof_c3.log 2>&1
ls -la $OUT | tail -5
