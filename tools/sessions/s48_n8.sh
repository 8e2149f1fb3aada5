#!/bin/bash
# round 2 (gpurun --gpus 8): the target configuration — C3 (1 M triangles, ptdirect 1920x1080, 1024 spp), STRONG scaling over the 8 GPUs
# of one box: bench.py under torchrun at N = 8 and 4 (product ncclReduce through ngi_gpu_comm_*), the nanogi CLI with --gpus 8
# (ngi_gpu_group_*: ncclCommInitAll, scene broadcast, one film reduce), C4 as BASELINE names it (4096 spp on 8 GPUs).
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/n8_gpus.txt
for n in 8 4; do
  NCCL_DEBUG=VERSION timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
      bench.py --gpus $n --steps 3 --warmup 3 > $OUT/n8_bench_c3_n$n.json 2> $OUT/n8_bench_c3_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > $OUT/n8_bench_c3_n1.json 2> $OUT/n8_bench_c3_n1.err
python - <<'PY'
import json
v = {}
for n in (1, 4, 8):
    try:
        j = json.loads([l for l in open(f"gpurun_out/n8_bench_c3_n{n}.json").read().splitlines() if l.startswith("{")][-1])
        v[n] = j
        print("N =", n, round(j["value"], 1), "Mpaths/s", j["scaling"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1), "film mean", j["film_mean"], "clocks", (j.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print("N =", n, "ERR", e)
for n in (4, 8):
    if 1 in v and n in v:
        print("strong-scaling efficiency at N = %d: %.3f (device-timed), %.3f (e2e)" % (n, v[n]["value"] / (n * v[1]["value"]), v[n]["e2e"]["value"] / (n * v[1]["e2e"]["value"])))
PY
python - <<'PY'
import os, sys, time
sys.path.insert(0, ".")
from nanogi_b200 import scenes
d = "gpurun_out/n8_c3_scene"; os.makedirs(d, exist_ok=True)
t0 = time.time(); print(scenes.write_scene_files(scenes.instanced_spheres(), d), round(time.time() - t0, 1), "s")
PY
for g in 1 8; do
  NCCL_DEBUG=VERSION timeout 900 nanogi_b200/nanogi ptdirect $OUT/n8_c3_scene/scene.yml $OUT/n8_c3_g$g.pfm 1920 1080 -n 2123366400 --seed 7 --gpus $g > $OUT/n8_cli_g$g.log 2>&1
  grep -E "Elapesed|GPU render|NCCL|BVH8" $OUT/n8_cli_g$g.log
done
rm -rf $OUT/n8_c3_scene $OUT/n8_c3_g1.pfm $OUT/n8_c3_g8.pfm
# C4 as BASELINE names it: 10 M triangles, 4096 spp over 8 GPUs (strong: 512 spp per GPU)
NCCL_DEBUG=VERSION timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
    bench.py --gpus 8 --workload c4 --steps 2 --warmup 3 --e2e-steps 1 > $OUT/n8_bench_c4_n8.json 2> $OUT/n8_bench_c4_n8.err
python - <<'PY'
import json
try:
    j = json.loads([l for l in open("gpurun_out/n8_bench_c4_n8.json").read().splitlines() if l.startswith("{")][-1])
    print("C4 N = 8", round(j["value"], 1), "Mpaths/s", round(j["mrays_per_s"]), "Mrays/s e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1))
except Exception as e:
    print("C4 ERR", e)
PY
