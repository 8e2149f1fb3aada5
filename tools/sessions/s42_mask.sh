#!/bin/bash
# round 2, session 42: table-driven node step (fixed triangle places) + fetch state in shared memory; node-group stack levels in shared
# memory (NGI_SSTACK 0 / 4 / 8); GPU tests; C4 sanity; C5 bit-exact check on C3.
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/s42_pytest.log 2>&1
tail -5 $OUT/s42_pytest.log
run() {  # tag lib workload spp env...
  tag=$1; lib=$2; wl=$3; spp=$4; shift 4
  env "$@" NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s42_${wl}_${tag}.json 2> $OUT/s42_${wl}_${tag}.err
  python - $OUT/s42_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s | extend ms", round(k["k_extend"]["avg_launch_ms"], 4), "shadow ms", round(k["k_shadow"]["avg_launch_ms"], 4),
          "logic ms", round(k[[x for x in k if x.startswith("logic")][0]]["avg_launch_ms"], 4), "| extend Grays/s", round(j["roofline"]["grays_per_s"], 3), "nodes", j["config"]["bvh8_nodes"], "build ms", round(j["config"]["bvh_build_ms"], 1),
          "scene MB", round(j["config"]["scene_device_bytes"] / 1e6, 1))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
P=$PWD
{
run ss0 $P/build/ss0.so c3 512 X=1
run ss4 $P/build/ss4.so c3 512 X=1
run ss8 $P/build/ss8.so c3 512 X=1
run ss0_t8 $P/build/ss0.so c3 512 NGI_TRACE_TRI_MIN=8
run ss0_t16 $P/build/ss0.so c3 512 NGI_TRACE_TRI_MIN=16
run ss8_t16 $P/build/ss8.so c3 512 NGI_TRACE_TRI_MIN=16
run ss0 $P/build/ss0.so c2 512 X=1
run ss4 $P/build/ss4.so c2 512 X=1
run ss8 $P/build/ss8.so c2 512 X=1
run ss0 $P/build/ss0.so c4 64 X=1
run ss8 $P/build/ss8.so c4 64 X=1
} | tee $OUT/s42_ab.txt
timeout 900 python tools/raybench.py --scene c3 --rays 16777216 --check 2097152 > $OUT/s42_raybench_c3.json 2> $OUT/s42_raybench_c3.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/s42_raybench_c3.json").read().strip().splitlines()[-1])
    print({k: (round(v["grays_per_s"], 2), v["checked"], v["mismatches"]) for k, v in j["batches"].items()})
except Exception as e:
    print("raybench ERR", e)
PY
rm -f $OUT/s42_iter.txt
NGI_LANES=1 NGI_ITER_LOG=$OUT/s42_iter.txt timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 \
    -f -o $OUT/s42_prof_c3 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 64 --no-cpu > $OUT/s42_prof_c3.log 2>&1
ls $OUT | grep s42 | wc -l
