#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 `pt` / `ptdirect` path (BASELINE.json metric: Mpaths/s, Mrays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c4pt] [--impl reference]

Own arm (default)
  step      one full render of the workload (one pass of the hot path over one batch of samples). The default workload
            is the configuration BASELINE.json's metric and target are quoted on — configs[2] = C3: procedural
            1M-triangle instanced-sphere scene, mixed diffuse / glossy / specular BSDFs, `ptdirect`, 1920 x 1080,
            1024 spp = 2 123 366 400 samples, unbounded path length (it fits one GPU; `--workload c2` is configs[1]).
            STRONG scaling: at any N the job is the same 2 123 366 400 samples; rank r renders the index range
            nanogi_b200.shard.shard_range gives it (disjoint Philox sample indices: the sample SET does not depend on
            N) and the per-rank films are summed by ONE NCCL reduce to rank 0 — the product's own ncclReduce
            (ngi_gpu_comm_reduce_film) — inside the timed region. (`--scaling weak` multiplies the job by N instead.)
  value     Mpaths/s of the whole job, device-timed (torch CUDA events on the stream the kernels are launched on),
            scene + BVH resident in HBM, film left in HBM; max over ranks.
  e2e       the same metric through the C-ABI call a host application makes, with HOST buffers, every step:
            ngi_gpu_scene_create (H2D of the flattened scene + GPU BVH build) -> render -> film D2H into pinned
            host memory -> ngi_gpu_scene_destroy; wall clock around the calls (they synchronise), max over ranks.
  roofline  dominant kernel (named in the JSON) timed live with CUDA events around every launch
            (NGI_RENDER_TIME_KERNELS pass on the same stream), algorithmic bytes per ray from SURVEY.md §8(d) against the
            measured HBM peak. The trace kernels are bound by instruction issue, not by HBM, on every BASELINE scene, so
            `bound` says "sm_issue" and the line carries issue-slot utilisation, active lanes per instruction, thread
            instructions per ray and the measured DRAM traffic from the committed `ncu --set full` capture of THIS build
            (profiles/r02_ncu_metrics.json, keyed by a hash of nanogi_b200/csrc; another build -> those fields are null).
  cpu_baseline  nanogi's own CPU code on all host cores (oracle/_ref: the reference's sources on stand-in third-party
            headers, Embree replaced by a scalar BVH; the oracle port when oracle/_ref is absent), bounded sample of the
            same workload; rank 0, N = 1 only.
Reference arm (`--impl reference`): the same CPU code on all host threads, same workload/metric, each step a bounded sample.

PyTorch is used for device memory, streams, events and torch.distributed only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpaths/s (ptdirect; Mrays/s and image checks alongside)"

DEFAULT_WORKLOAD = "c3"     # the configuration BASELINE.json's metric / target is stated on (fits one GPU)

WORKLOADS = {
    # name: (scene generator, renderer, W, H, spp, max_num_vertices, description)
    "c1": ("cornell_box", "pt", 256, 256, 64, 8, "C1 Cornell box (38 tris) pt 256x256 64spp -m 8"),
    "c2": ("cornell_spheres", "ptdirect", 1024, 1024, 1024, -1,
           "C2 Cornell box + glossy/glass icospheres (2598 tris) ptdirect 1024x1024 1024spp"),
    "c3": ("instanced_spheres", "ptdirect", 1920, 1080, 1024, -1,
           "C3 procedural 1M-triangle instanced spheres, mixed D/G/S, ptdirect 1920x1080 1024spp"),
    "c4": ("interior", "ptdirect", 1920, 1080, 4096, -1,
           "C4 procedural 10M-triangle interior (Sibenik-like layout, 256 small area lights), ptdirect 1920x1080 4096spp"),
    "c4pt": ("interior", "pt", 1920, 1080, 4096, -1,
             "C4 procedural 10M-triangle interior (Sibenik-like layout, 256 small area lights), pt 1920x1080 4096spp"),
    # the wavefront bdpt (not a BASELINE config; same contract, so that its numbers are measured the same way)
    "c2bdpt": ("cornell_spheres", "bdpt", 1024, 1024, 256, -1,
               "C2's scene (Cornell box + glossy/glass icospheres, 2598 tris) bdpt 1024x1024 256spp"),
}


def b_ray(n_tris: int) -> int:
    """Algorithmic bytes per ray, SURVEY.md §8(d): d*80 + 4*48 + 96 with d = ceil(log8(nTri/4))."""
    d, cap = 0, 4
    while cap < n_tris:
        cap *= 8
        d += 1
    return max(d, 1) * 80 + 4 * 48 + 96


def csrc_sha() -> str:
    """Hash of the CUDA sources the loaded module was built from (ties profiles/r02_ncu_metrics.json to a build)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "nanogi_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".h", ".cuh", ".cu")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_metrics(workload: str, kernel: str):
    """Counters of ONE steady-state launch of `kernel` from the committed `ncu --set full` capture (tools/ncu_metrics.py),
    or None when the capture is of another build of the CUDA sources."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    try:
        j = json.load(open(p))
    except Exception:
        return None
    if j.get("csrc_sha") != csrc_sha():
        return None
    return (j.get("workloads", {}).get(workload, {}) or {}).get(kernel)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            for key in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if key in j:
                    return float(j[key]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def build_scene(name: str, aspect: float):
    from nanogi_b200 import scenes
    gen = getattr(scenes, WORKLOADS[name][0])
    return scenes.to_scene_data(gen(), aspect, name=name)


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's CPU implementation of the path on the host cores. kind "reference": the reference's OWN sources
    (src/nanogi.cpp + include/nanogi/*.hpp) built by oracle/build_ref.sh against the stand-in libraries of oracle/refshim —
    its TBB-style sample loop, Primitive functions and film handling, with Embree's kernels replaced by a scalar BVH (so its
    ray queries are slower than real Embree's SIMD ones). kind "port": the oracle restatement, when oracle/_ref is not there."""

    def __init__(self, workload: str, sd, aspect: float):
        from nanogi_b200 import scenes
        from oracle import pyoracle, pyref
        self.cores = os.cpu_count() or 1
        self.orc = pyoracle.OracleScene(sd)
        self.ref = None
        # (C3's 983 k triangles load in ~25 s through the reference's loader; C4's 10 M would take minutes: the port runs there)
        if pyref.available() and sd.num_tris <= 2000000 and not os.environ.get("NGI_BENCH_CPU_PORT"):
            # the reference loads scene FILES: write the workload as schema.yml + OBJ (not part of any timed region)
            self.ref = pyref.RefScene(getattr(scenes, WORKLOADS[workload][0])(), aspect)
        self.kind = "reference" if self.ref is not None else "port"
        self.note = ("nanogi's own sources (oracle/_ref: src/nanogi.cpp + include/nanogi/*.hpp on stand-in third-party headers; Embree's "
                     "kernels replaced by a scalar BVH)" if self.ref is not None else
                     "restated nanogi CPU path (oracle port, own BVH, no Embree)")

    def render(self, renderer, n, W, H, m, seed):
        """returns seconds"""
        t0 = time.perf_counter()
        if self.ref is not None:
            self.ref.render(renderer, n, W, H, max_num_vertices=m, seed=seed, num_threads=self.cores)
        else:
            self.orc.render(renderer, n, W, H, max_num_vertices=m, seed=seed)
        return time.perf_counter() - t0

    def rays_per_path(self, renderer, W, H, m):
        _, st = self.orc.render(renderer, 1 << 18, W, H, max_num_vertices=m, seed=3)
        return (st["extend_rays"] + st["shadow_rays"]) / float(1 << 18)


def run_reference(args, rank: int):
    """The reference arm: nanogi's own CPU implementation of the path on all host threads (see CpuReference)."""
    if rank != 0:
        return
    gen, renderer, W, H, spp, m, desc = WORKLOADS[args.workload]
    sd = build_scene(args.workload, W / H)
    cpu = CpuReference(args.workload, sd, W / H)
    cores = cpu.cores
    # bounded sample: calibrate on 2^18 samples, then size a step to ~ args.cpu_step_seconds of wall time
    n_step = 1 << 18
    while True:
        dt0 = cpu.render(renderer, n_step, W, H, m, 11)
        if dt0 >= 0.6 * args.cpu_step_seconds or n_step >= W * H * spp:
            break
        n_step = int(min(W * H * spp, n_step * min(8.0, max(1.5, args.cpu_step_seconds / max(dt0, 1e-3)))))
    for i in range(args.warmup):
        cpu.render(renderer, n_step, W, H, m, 100 + i)
    dt = 0.0
    for i in range(args.steps):
        dt += cpu.render(renderer, n_step, W, H, m, 200 + i)
    rays = cpu.rays_per_path(renderer, W, H, m) * n_step * args.steps
    v = n_step * args.steps / dt / 1e6
    sample = f"{n_step} samples/step ({n_step / (W * H):.2f} spp of {spp}) x {args.steps} steps, same scene/resolution/renderer"
    line = {
        "impl": "reference", "metric": METRIC.replace("ptdirect", renderer), "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "mrays_per_s": rays / dt / 1e6,
        "config": {"workload": desc, "renderer": renderer, "width": W, "height": H, "spp": spp, "samples_per_step": W * H * spp, "max_num_vertices": m,
                   "sampled": f"each step renders {n_step} samples of the job (bounded CPU sample)",
                   "note": cpu.note},
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_own(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from nanogi_b200 import capi, shard

    if not torch.cuda.is_available() or capi.device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device — the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    gen, renderer, W, H, spp, m, desc = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    # STRONG scaling (default): the job is the workload's W*H*spp samples at any N, rank r takes the index range shard_range gives
    # it (src/nanogi.cpp:281: one parallel_for over [0, NumSamples)). --scaling weak: N times the job, fixed work per GPU.
    n_total = W * H * spp * (world if args.scaling == "weak" else 1)
    sample_offset, n_rank = shard.shard_range(n_total, rank, world)
    comm = shard.make_comm(local_rank)         # the product's NCCL communicator (ncclCommInitRank; id broadcast over the process group)
    sd = build_scene(args.workload, W / H)
    scene = capi.GpuScene(sd, local_rank)
    info = scene.info()
    film = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    # ONE explicit stream for everything: the module forks its lanes from it and joins them back, the product's ncclReduce runs on it,
    # the timing events are recorded on it. (With torch's default stream — handle 0 — the module would render on its own non-blocking
    # stream and the reduce on the legacy default stream: nothing orders the next step's film reset after a reduce still in flight.
    # At N = 8 that race added 3 of the 8 ring chunks of the previous image to the next one, s48.)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(i: int, flags: int = 0):
        flush.zero_()                                               # L2 flush between iterations
        st = scene.render_device(film.data_ptr(), stream.cuda_stream, renderer, n_rank, W, H, max_num_vertices=m,
                                 seed=1000 + i, sample_offset=sample_offset, film_norm_samples=n_total, flags=flags,
                                 wave_capacity=args.wave_capacity)
        shard.reduce_film(film, 0, comm, stream.cuda_stream)        # the one exchange of the path (src/nanogi.cpp:429-437): ncclReduce
        return st

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = ext = sh = iters = 0
    ev0.record(stream)
    for i in range(args.steps):
        st = step(args.warmup + i)
        launches += st.kernel_launches; ext += st.extend_rays; sh += st.shadow_rays; iters += st.wave_iterations
    ev1.record(stream)
    barrier()
    clk = clocks.stop() if clocks else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(launches), float(ext), float(sh)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t = float(ms.item()) * 1e-3
    value = n_total * args.steps / t / 1e6
    mrays = float(cnt[1] + cnt[2]) / t / 1e6
    film_mean = float(film.mean().item())

    # ---- end to end through the C ABI with host buffers -------------------------------------------------
    pinned = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
    h2d = sd.positions.nbytes + sd.normals.nbytes + len(sd.prims) * __import__("ctypes").sizeof(capi.NgiPrimitive) \
        + __import__("ctypes").sizeof(capi.NgiRenderParams)
    d2h = film.numel() * 4

    e2e_parts = {"create_s": 0.0, "render_s": 0.0, "destroy_s": 0.0}

    def e2e_step(i: int):
        ta = time.perf_counter()
        sc = capi.GpuScene(sd, local_rank)                           # H2D of the scene + GPU BVH build
        tb = time.perf_counter()
        if world == 1:
            p = sc._params(renderer, n_rank, W, H, max_num_vertices=m, seed=5000 + i, film_norm_samples=n_total,
                           wave_capacity=args.wave_capacity)
            st = capi.NgiRenderStats()
            capi._check(sc.lib.ngi_gpu_render(sc.handle, __import__("ctypes").byref(p), pinned.data_ptr(), __import__("ctypes").byref(st)),
                        "ngi_gpu_render")                           # render + film D2H into pinned host memory
        else:
            sc.render_device(film.data_ptr(), stream.cuda_stream, renderer, n_rank, W, H, max_num_vertices=m, seed=5000 + i,
                             sample_offset=sample_offset, film_norm_samples=n_total, wave_capacity=args.wave_capacity)
            shard.reduce_film(film, 0, comm, stream.cuda_stream)
            if rank == 0:
                pinned.copy_(film, non_blocking=True)
            torch.cuda.synchronize(dev)
        tc = time.perf_counter()
        sc.close()
        td = time.perf_counter()
        if i >= 0:
            e2e_parts["create_s"] += tb - ta; e2e_parts["render_s"] += tc - tb; e2e_parts["destroy_s"] += td - tc

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step(-1)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(te.item()) / 1e6

    # ---- roofline of the dominant kernel: per-launch CUDA events on the same stream ------------------------
    roof = None
    kern = None
    if rank == 0:
        stt = scene.render_device(film.data_ptr(), stream.cuda_stream, renderer, n_rank, W, H, max_num_vertices=m, seed=77,
                                  sample_offset=0, film_norm_samples=n_rank, flags=capi.RENDER_TIME_KERNELS,
                                  wave_capacity=args.wave_capacity)
        peak, peak_src = measured_peaks()
        br = b_ray(int(info.num_tris))
        bd = renderer == "bdpt"
        k_ext, k_sh = ("k_bdw_extend", "k_bdw_shadow") if bd else ("k_extend", "k_shadow")
        k_logic = "logic per batch (k_bdw_start + k_bdw_step + k_bdw_count + k_bdw_expand + sort + k_bdw_contrib)" if bd else "logic (k_classify + k_surface + k_eye)"
        per = {k_logic: (stt.logic_kernel_seconds, stt.logic_launches, None),
               k_ext: (stt.extend_kernel_seconds, stt.extend_launches, stt.extend_rays),
               k_sh: (stt.shadow_kernel_seconds, stt.shadow_launches, stt.shadow_rays)}
        total_k = sum(v[0] for v in per.values()) or 1.0
        kern = {k: {"seconds": v[0], "launches": int(v[1]), "share": v[0] / total_k,
                    "avg_launch_ms": (v[0] / v[1] * 1e3 if v[1] else None)} for k, v in per.items()}
        # dominant TRACE kernel (the path's bytes are the BVH/triangle fetches of the ray queries)
        dom = k_ext if stt.extend_kernel_seconds >= stt.shadow_kernel_seconds else k_sh
        sec, nl, rays = per[dom]
        achieved = rays * br / sec / 1e9 if sec > 0 else 0.0
        rays_per_launch = rays / max(nl, 1)
        # counters of one steady-state launch of the same kernel from the committed ncu capture of THIS build
        nm = ncu_metrics(args.workload, dom)
        traffic = traffic_src = None
        counters = None
        if nm:
            per_ray = nm["dram_bytes"] / nm["rays"]
            traffic = per_ray * rays_per_launch                        # bytes per launch of THIS run's size
            traffic_src = "%s: %.0f B/ray of DRAM traffic (x %.0f rays per launch here) vs %d algorithmic B/ray" % (nm["source"], per_ray, rays_per_launch, br)
            counters = {k: nm.get(k) for k in ("issue_active_pct", "active_lanes", "thread_inst_per_ray", "warp_inst_per_ray", "alu_pipe_pct", "fma_pipe_pct",
                                               "dram_throughput_pct", "l1_hit_pct", "l2_hit_pct", "warps_active_pct", "registers", "local_load_inst_per_ray",
                                               "local_store_inst_per_ray")}
        roof = {"bound": "sm_issue", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_ray": br, "rays_per_launch": rays_per_launch,
                "avg_launch_ms": sec / max(nl, 1) * 1e3, "grays_per_s": rays / sec / 1e9 if sec > 0 else 0.0,
                "ncu": counters,
                "note": "achieved = algorithmic bytes (SURVEY 8d: %d B/ray) x rays per launch / CUDA-event launch time, against the measured HBM "
                        "copy peak: an ALGORITHMIC fraction. The kernel is bound by instruction issue (BVH8 node steps + triangle tests; the "
                        "BVH, %.1f MB of nodes + triangles, is served from L1/L2), so `ncu` carries the physical counters of one steady-state "
                        "launch of this build: issue-slot utilisation, active lanes, thread instructions per ray, measured DRAM bytes"
                        % (br, (int(info.bvh8_nodes) * 80 + int(info.num_tris) * 48) / 1e6)}

    # ---- CPU baseline (oracle port) on the host cores: rank 0, N = 1 only -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ref = CpuReference(args.workload, sd, W / H)
        # bounded sample: grow it until one render takes about --cpu-seconds (a short render over-states the rate: thread start-up
        # and the per-thread films are paid per call)
        n_cpu, dtc = 1 << 18, 0.0
        while True:
            dtc = ref.render(renderer, n_cpu, W, H, m, 12)
            if dtc >= 0.6 * args.cpu_seconds or n_cpu >= n_rank:
                break
            n_cpu = int(min(n_rank, n_cpu * min(8.0, max(1.5, args.cpu_seconds / max(dtc, 1e-3)))))
        cpu = {"value": n_cpu / dtc / 1e6, "unit": "Mpaths/s", "cores": ref.cores, "kind": ref.kind,
               "sample": f"{n_cpu} samples ({n_cpu / (W * H):.2f} spp of {spp}) of the same scene/resolution/renderer, {dtc:.1f} s",
               "mrays_per_s": ref.rays_per_path(renderer, W, H, m) * n_cpu / dtc / 1e6, "note": ref.note}
        if ref.kind == "reference":          # the oracle port next to it, for scale (its ray queries use a SAH BVH)
            t0 = time.perf_counter()
            ref.orc.render(renderer, n_cpu, W, H, max_num_vertices=m, seed=12)
            cpu["oracle_port_value"] = n_cpu / (time.perf_counter() - t0) / 1e6

    # the module's default slot count per lane (render_impl in ngi_gpu.cu)
    wave_slots = args.wave_capacity or ((1 << 23) if n_rank >= (1 << 30) else (1 << 22) if n_rank >= (1 << 27) else (1 << 21))
    if rank == 0:
        line = {
            "metric": METRIC.replace("ptdirect", renderer), "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mrays_per_s": mrays,
            "config": {"workload": desc, "renderer": renderer, "width": W, "height": H, "spp": n_total // (W * H), "samples_per_step": n_total,
                       "samples_per_gpu": n_rank,
                       "max_num_vertices": m, "tris": int(info.num_tris), "bvh8_nodes": int(info.bvh8_nodes),
                       "scene_device_bytes": int(info.device_bytes), "bvh_build_ms": info.build_gpu_seconds * 1e3,
                       "parallelism": f"one job of {n_total} samples sharded by index over {world} GPU(s) ({args.scaling} scaling); scene replicated; "
                                      "one ncclReduce of the films inside the timed region (libnanogi_gpu.so: ngi_gpu_comm_reduce_film)",
                       "l2": ("256 MB memset between iterations flushes L2; a bdpt batch (vertex + cache records of 2 subpaths per sample, "
                              "GBs per batch, two in flight) also exceeds it") if renderer == "bdpt" else
                             "256 MB memset between iterations flushes L2; wavefront state (%.0f MB: 188 B x %d slots x 2 lanes) also exceeds it"
                             % (wave_slots * 188 * 2 / 1e6, wave_slots)},
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "includes": "scene H2D + GPU BVH build + render + film D2H (pinned) + destroy, wall clock",
                    "breakdown_s_per_step": {k: v / e2e_steps for k, v in e2e_parts.items()}},
            "gpu_launches": int(cnt[0].item()), "wave_iterations": int(iters), "film_mean": film_mean,
            "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "clocks": clk,
        }
        emit(line)
    scene.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict):
    """rank 0 prints ONE JSON line on the process's ORIGINAL stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # rank 0 prints ONE JSON line on stdout. NCCL writes its version banner ("NCCL version 2.28.9+cuda12.9") to the C stdout
    # of rank 0 at init whenever NCCL_DEBUG >= VERSION: keep fd 1 for the JSON line only and send everything else to stderr.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the same job at any N; weak: N times the job")
    ap.add_argument("--spp", type=int, default=0, help="override the samples per pixel of the job (default: the workload's)")
    ap.add_argument("--wave-capacity", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="size of the cpu_baseline sample")
    ap.add_argument("--cpu-step-seconds", type=float, default=4.0, help="--impl reference: CPU seconds per step")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        print("bench.py: note: W < 3 warm-up steps — not a valid headline number", file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd, stdout=_JSON_FD))     # the ranks inherit the ORIGINAL stdout
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    run_own(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
