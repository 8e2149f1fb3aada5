#!/bin/bash
# round 2, session 43: several node steps per round of the trace loop (NGI_NODE_REPS) x backlog size; GPU tests incl. the new
# branch tests and the gated image acceptance; the default bench line with the ncu-metrics JSON in place.
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag lib workload spp env...
  tag=$1; lib=$2; wl=$3; spp=$4; shift 4
  env "$@" NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s43_${wl}_${tag}.json 2> $OUT/s43_${wl}_${tag}.err
  python - $OUT/s43_${wl}_${tag}.json <<'PY'
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
for v in r1 r2_tq2 r2_tq3 r2_tq4 r3_tq4; do run $v $P/build/$v.so c3 512 X=1; done
run r2_tq4_t16 $P/build/r2_tq4.so c3 512 NGI_TRACE_TRI_MIN=16
run r2_tq4_t20 $P/build/r2_tq4.so c3 512 NGI_TRACE_TRI_MIN=20
run r3_tq4_t20 $P/build/r3_tq4.so c3 512 NGI_TRACE_TRI_MIN=20
for v in r1 r2_tq4 r3_tq4; do run $v $P/build/$v.so c2 512 X=1; done
for v in r1 r2_tq4; do run $v $P/build/$v.so c4 64 X=1; done
} | tee $OUT/s43_ab.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > $OUT/s43_pytest.log 2>&1
tail -25 $OUT/s43_pytest.log
timeout 600 python bench.py --no-cpu > $OUT/s43_bench_default.json 2> $OUT/s43_bench_default.err
python - <<'PY'
import json
try:
    j = json.loads([l for l in open("gpurun_out/s43_bench_default.json").read().splitlines() if l.startswith("{")][-1])
    print(round(j["value"], 2), j["scaling"], "e2e", round(j["e2e"]["value"], 2), json.dumps(j["roofline"])[:900])
except Exception as e:
    print("ERR", e)
PY
