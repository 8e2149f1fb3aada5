#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the wavefront kernels in steady state.
# usage: tools/gpu_profile.sh <tag> [workload] [spp] [skip]
TAG=${1:-r1}; WL=${2:-c2}; SPP=${3:-64}; SKIP=${4:-60}
OUT=gpurun_out
mkdir -p $OUT
# every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload $WL --spp 8 --no-cpu > $OUT/launches_${TAG}.log 2>&1
# full capture of the three wavefront kernels, taken well into the steady state of the first render
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_classify|k_surface|k_eye' -s $SKIP -c 5 \
    -f -o $OUT/prof_${TAG} python bench.py --steps 1 --warmup 1 --e2e-steps 1 --workload $WL --spp $SPP --no-cpu > $OUT/prof_${TAG}.log 2>&1
ls -la $OUT
