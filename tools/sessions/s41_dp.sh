#!/bin/bash
# round 2, session 41: GPU tests at HEAD (film fold, group API, new CLI); SAH-optimal collapse vs greedy; TQ sizes / tri_min; FFMA2;
# the new default bench line (C3, strong scaling) + reference arm; ncu capture with exact per-launch ray counts.
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/s41_pytest.log 2>&1
tail -5 $OUT/s41_pytest.log
run() {  # tag lib workload env...
  tag=$1; lib=$2; wl=$3; shift 3
  env "$@" NGI_GPU_LIB=$lib timeout 300 python bench.py --workload $wl --spp 512 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s41_${wl}_${tag}.json 2> $OUT/s41_${wl}_${tag}.err
  python - $OUT/s41_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s | extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4),
          "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4), "| extend Grays/s", round(j["roofline"]["grays_per_s"], 3), "nodes", j["config"]["bvh8_nodes"], "build ms", round(j["config"]["bvh_build_ms"], 1))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
P=$PWD
{
run old_greedy $P/build/old.so c3 NGI_COLLAPSE_GREEDY=1
run dflt       $P/nanogi_b200/libnanogi_gpu.so c3 X=1
run greedy     $P/nanogi_b200/libnanogi_gpu.so c3 NGI_COLLAPSE_GREEDY=1
run cp02       $P/nanogi_b200/libnanogi_gpu.so c3 NGI_SAH_CPRIM=0.2
run cp05       $P/nanogi_b200/libnanogi_gpu.so c3 NGI_SAH_CPRIM=0.5
run t12        $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=12
run t16        $P/nanogi_b200/libnanogi_gpu.so c3 NGI_TRACE_TRI_MIN=16
run tq4_t12    $P/build/tq4.so c3 NGI_TRACE_TRI_MIN=12
run ffma2_t12  $P/build/ffma2.so c3 NGI_TRACE_TRI_MIN=12
run old_greedy $P/build/old.so c2 NGI_COLLAPSE_GREEDY=1
run dflt       $P/nanogi_b200/libnanogi_gpu.so c2 X=1
run t12        $P/nanogi_b200/libnanogi_gpu.so c2 NGI_TRACE_TRI_MIN=12
} | tee $OUT/s41_ab.txt
timeout 600 python bench.py > $OUT/s41_bench_default.json 2> $OUT/s41_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/s41_bench_ref.json 2> $OUT/s41_bench_ref.err
tail -c 600 $OUT/s41_bench_default.err; tail -c 300 $OUT/s41_bench_ref.err
rm -f $OUT/s41_iter.txt
NGI_LANES=1 NGI_ITER_LOG=$OUT/s41_iter.txt timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
    -f -o $OUT/s41_prof_c3 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 64 --no-cpu > $OUT/s41_prof_c3.log 2>&1
python - <<'PY'
import json
for f in ("s41_bench_default", "s41_bench_ref"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, j["config"]["workload"][:30], round(j["value"], 2), j["scaling"], "e2e", round(j["e2e"]["value"], 2), "cpu", (j.get("cpu_baseline") or {}).get("value"), (j.get("cpu_baseline") or {}).get("kind"), "frac", (j.get("roofline") or {}).get("frac"), (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls $OUT | grep s41 | wc -l
