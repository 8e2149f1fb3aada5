#!/bin/bash
# round 2, session 45: single shared-memory pool with a per-thread base register (inline PTX ld/st.shared) vs __shared__ arrays;
# full GPU suite; ncu capture + metrics; default bench line.
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag lib workload spp env...
  tag=$1; lib=$2; wl=$3; spp=$4; shift 4
  env "$@" NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s45_${wl}_${tag}.json 2> $OUT/s45_${wl}_${tag}.err
  python - $OUT/s45_${wl}_${tag}.json <<'PY'
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
run arrays $P/build/r2_tq2.so c3 512 X=1
run pool   $P/build/pool.so c3 512 X=1
run pool_t8   $P/build/pool.so c3 512 NGI_TRACE_TRI_MIN=8
run pool_t16  $P/build/pool.so c3 512 NGI_TRACE_TRI_MIN=16
run pool_r2   $P/build/pool.so c3 512 NGI_TRACE_REFILL_MIN=2
run pool_r8   $P/build/pool.so c3 512 NGI_TRACE_REFILL_MIN=8
run pool_c32  $P/build/pool.so c3 512 NGI_TRACE_CHUNK=32
run pool_c128 $P/build/pool.so c3 512 NGI_TRACE_CHUNK=128
run arrays $P/build/r2_tq2.so c2 512 X=1
run pool   $P/build/pool.so c2 512 X=1
run pool   $P/build/pool.so c4 64 X=1
} | tee $OUT/s45_ab.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/s45_pytest.log 2>&1
tail -6 $OUT/s45_pytest.log
rm -f $OUT/s45_iter.txt
NGI_LANES=1 NGI_ITER_LOG=$OUT/s45_iter.txt timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
    -f -o $OUT/s45_prof_c3 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 64 --no-cpu > $OUT/s45_prof_c3.log 2>&1
timeout 600 python bench.py > $OUT/s45_bench_default.json 2> $OUT/s45_bench_default.err
python - <<'PY'
import json
try:
    j = json.loads([l for l in open("gpurun_out/s45_bench_default.json").read().splitlines() if l.startswith("{")][-1])
    print(round(j["value"], 2), j["scaling"], "e2e", round(j["e2e"]["value"], 2), "cpu", j["cpu_baseline"], "frac", j["roofline"]["frac"], j["clocks"])
except Exception as e:
    print("ERR", e)
PY
