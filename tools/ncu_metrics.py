#!/usr/bin/env python
"""Per-ray counters of the trace kernels from one `ncu --set full` capture -> profiles/r02_ncu_metrics.json (read by bench.py).

    python tools/ncu_metrics.py <workload> <capture.ncu-rep> <iter_log.txt> <skip> [--out profiles/r02_ncu_metrics.json]

The capture is taken by tools/sessions/*: `NGI_LANES=1 NGI_ITER_LOG=<iter_log> ncu --set full -k regex:'k_extend|k_shadow' -s <skip> -c 4 ...
python bench.py --workload <wl> ...`. With one lane the launches alternate k_extend / k_shadow, one pair per wavefront iteration, so the
j-th captured launch of a kernel belongs to iteration skip/2 + j of the FIRST render of the process, and the NGI_ITER_LOG file (written by
libnanogi_gpu.so, ngi_gpu.cu) gives the exact number of rays that launch traced. The JSON carries a hash of nanogi_b200/csrc so that
bench.py only quotes counters measured on the build it is running.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def scaled(v, unit):
    """ncu prints byte counters with a unit column (byte, Kbyte, Mbyte, Gbyte)."""
    if v is None:
        return None
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("rep")
    ap.add_argument("iter_log")
    ap.add_argument("skip", type=int)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_ncu_metrics.json"))
    args = ap.parse_args()
    import bench

    # rays per iteration of lane 0 of the first render in the log
    iters = {}
    first = True
    for ln in open(args.iter_log):
        if ln.startswith("#"):
            if iters:
                break
            continue
        k, it, sh, ex = (int(x) for x in ln.split())
        if k == 0:
            iters[it] = {"k_shadow": sh, "k_extend": ex}
    hdr, units, rows = raw(args.rep)
    ix = {h: i for i, h in enumerate(hdr)}

    def get(r, key, unit_scaled=False):
        if key not in ix:
            return None
        v = num(r[ix[key]])
        return scaled(v, units[ix[key]]) if unit_scaled else v

    def per_ray(r, key, rays):
        v = get(r, key)
        return None if v is None else v / rays

    def time_us(r):
        v = get(r, "gpu__time_duration.sum")
        return None if v is None else v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1.0)

    seen = {"k_extend": 0, "k_shadow": 0}
    per = {}
    for r in rows:
        name = r[ix["Kernel Name"]]
        kern = "k_extend" if "k_extend" in name else "k_shadow" if "k_shadow" in name else None
        if kern is None:
            continue
        it = args.skip // 2 + seen[kern]
        seen[kern] += 1
        rays = iters.get(it, {}).get(kern)
        if not rays:
            continue
        dram = (get(r, "dram__bytes_read.sum", True) or 0) + (get(r, "dram__bytes_write.sum", True) or 0)
        winst = get(r, "smsp__inst_executed.sum")
        lanes = get(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
        m = {
            "source": os.path.basename(args.rep), "iteration": it, "rays": rays, "duration_us": time_us(r),
            "dram_bytes": dram, "dram_bytes_per_ray": dram / rays,
            "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "active_lanes": lanes,
            "warp_inst_per_ray": winst / rays if winst else None,
            "thread_inst_per_ray": winst * lanes / rays if winst and lanes else None,
            "alu_pipe_pct": get(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": get(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1_hit_pct": get(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct"),
            "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": get(r, "launch__registers_per_thread"),
            # the traversal stack (node groups) is thread-local memory, the triangle backlog shared memory: warp instructions per ray
            "local_load_inst_per_ray": per_ray(r, "smsp__sass_inst_executed_op_local_ld.sum", rays),
            "local_store_inst_per_ray": per_ray(r, "smsp__sass_inst_executed_op_local_st.sum", rays),
            "local_load_l1_hit_pct": get(r, "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"),
            "shared_load_inst_per_ray": per_ray(r, "smsp__sass_inst_executed_op_shared_ld.sum", rays),
            "shared_store_inst_per_ray": per_ray(r, "smsp__sass_inst_executed_op_shared_st.sum", rays),
        }
        per.setdefault(kern, m)          # the first captured launch of each kernel
    j = {}
    if os.path.exists(args.out):
        try:
            j = json.load(open(args.out))
        except Exception:
            j = {}
    sha = bench.csrc_sha()
    if j.get("csrc_sha") != sha:
        j = {"csrc_sha": sha, "workloads": {}}
    j["note"] = ("one steady-state launch per kernel from `ncu --set full --clock-control none` (tools/ncu_metrics.py); rays = exact count of that "
                 "launch (NGI_ITER_LOG); csrc_sha = sha256 over nanogi_b200/csrc at capture time")
    j["workloads"][args.workload] = per
    json.dump(j, open(args.out, "w"), indent=1)
    print(json.dumps(per, indent=1))


if __name__ == "__main__":
    main()
