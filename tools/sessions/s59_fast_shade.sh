#!/bin/bash
# round 2, session 59 (1 GPU): SFU division / sqrt / sincos in the shading code (NGI_FAST_SHADE, in-tree build) against HEAD's module
# (build/head.so): GPU suite on the new build, A/B bench lines on C3 / C2 / C4, occupancy variants of k_surface / k_eye on top of it,
# ncu captures of the logic kernels before / after.
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/s59_pytest.log 2>&1
tail -5 $OUT/s59_pytest.log
run() {  # tag lib workload spp
  tag=$1; lib=$2; wl=$3; spp=$4
  NGI_GPU_LIB=$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --e2e-steps 1 --no-cpu > $OUT/s59_${wl}_${tag}.json 2> $OUT/s59_${wl}_${tag}.err
  python - $OUT/s59_${wl}_${tag}.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k = j.get("kernels") or {}
    print(sys.argv[1], round(j["value"], 1), "Mpaths/s e2e", round(j["e2e"]["value"], 1), "|",
          {n.split(" ")[0]: (round(v["avg_launch_ms"], 4), round(v.get("share", 0), 3)) for n, v in k.items()}, "| film mean", j.get("film_mean"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
P=$PWD
{
run head $P/build/head.so c3 512
run fast $P/nanogi_b200/libnanogi_gpu.so c3 512
run fast_s5 $P/build/fast_s5.so c3 512
run fast_s3 $P/build/fast_s3.so c3 512
run fast_e5 $P/build/fast_e5.so c3 512
run head $P/build/head.so c2 512
run fast $P/nanogi_b200/libnanogi_gpu.so c2 512
run fast_s5 $P/build/fast_s5.so c2 512
run fast_s3 $P/build/fast_s3.so c2 512
run fast_e5 $P/build/fast_e5.so c2 512
run head $P/build/head.so c1 4096
run fast $P/nanogi_b200/libnanogi_gpu.so c1 4096
run head $P/build/head.so c2bdpt 64
run fast $P/nanogi_b200/libnanogi_gpu.so c2bdpt 64
run head $P/build/head.so c4 64
run fast $P/nanogi_b200/libnanogi_gpu.so c4 64
} | tee $OUT/s59_ab.txt
for v in head fast; do
  lib=$P/nanogi_b200/libnanogi_gpu.so; [ $v = head ] && lib=$P/build/head.so
  NGI_GPU_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_surface|k_eye' -s 90 -c 3 \
      -f -o $OUT/s59_prof_c2_logic_$v python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c2 --spp 64 --no-cpu > $OUT/s59_prof_c2_logic_$v.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_surface|k_eye' -s 90 -c 3 \
    -f -o $OUT/s59_prof_c3_logic_fast python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload c3 --spp 64 --no-cpu > $OUT/s59_prof_c3_logic_fast.log 2>&1
ls -la $OUT | grep s59 | wc -l
