#!/bin/bash
# round 2, session 44 (gpurun --gpus 2): the product's NCCL paths on two GPUs — ngi_gpu_group_* (ncclCommInitAll, scene broadcast, film
# reduce) through the tests and the nanogi CLI, ngi_gpu_comm_* (ncclCommInitRank) through bench.py under torchrun; strong scaling 1 -> 2.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/s44_gpus.txt
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q ) > $OUT/s44_pytest.log 2>&1
tail -6 $OUT/s44_pytest.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/s44_bench_n2.json 2> $OUT/s44_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu > $OUT/s44_bench_n1.json 2> $OUT/s44_bench_n1.err
python - <<'PY'
import json
v = {}
for n in (1, 2):
    try:
        j = json.loads([l for l in open(f"gpurun_out/s44_bench_n{n}.json").read().splitlines() if l.startswith("{")][-1])
        v[n] = j
        print("N =", n, round(j["value"], 1), "Mpaths/s", j["scaling"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1), "film mean", j["film_mean"], "launches", j["gpu_launches"])
    except Exception as e:
        print("N =", n, "ERR", e)
if 1 in v and 2 in v:
    print("strong-scaling efficiency at N = 2: %.3f (device-timed), %.3f (e2e)" % (v[2]["value"] / (2 * v[1]["value"]), v[2]["e2e"]["value"] / (2 * v[1]["e2e"]["value"])))
PY
grep -E "NCCL INFO (Connected|comm|ncclComm|Using|NVLS)|nanogi_gpu\] NCCL" $OUT/s44_bench_n2.err | head -20
# the drop-in CLI on C3: scene files written once, then 1 GPU vs 2 GPUs
python - <<'PY'
import os, sys, time
sys.path.insert(0, ".")
from nanogi_b200 import scenes
d = "gpurun_out/s44_c3_scene"; os.makedirs(d, exist_ok=True)
t0 = time.time(); print(scenes.write_scene_files(scenes.instanced_spheres(), d), round(time.time() - t0, 1), "s")
PY
for g in 1 2; do
  NCCL_DEBUG=VERSION timeout 900 nanogi_b200/nanogi ptdirect $OUT/s44_c3_scene/scene.yml $OUT/s44_c3_g$g.pfm 1920 1080 -n 2123366400 --seed 7 --gpus $g > $OUT/s44_cli_g$g.log 2>&1
  grep -E "Elapesed|GPU render|NCCL|BVH8" $OUT/s44_cli_g$g.log
done
rm -rf $OUT/s44_c3_scene $OUT/s44_c3_g1.pfm $OUT/s44_c3_g2.pfm
