#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_<tag>.csv  > profiles/<round>_launches_<tag>.txt
    python tools/ncu_summary.py full     gpurun_out/prof_<tag>.ncu-rep  > profiles/<round>_ncu_<tag>.txt
    python tools/ncu_summary.py source   gpurun_out/prof_<tag>.ncu-rep <kernel-id> [top]   (hot SASS/source lines)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name) if name.startswith("void ") else name
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {path}")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:58]:58s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:9.1f} {v[1] / tot * 100:6.1f}%")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def full(path):
    hdr, units, rows = raw(path)
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none --import-source on : {path}")
    for r in rows:
        print(f"\n== launch id {r[idx['ID']]}: {r[idx['Kernel Name']][:80]}")
        for k in KEYS:
            if k in idx:
                print(f"  {k:82s} {r[idx[k]]:>18s} {units[idx[k]]}")


def source(path, kid, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-kernel-base", "function"], capture_output=True, text=True).stdout
    blocks = out.split("\n\n")
    print(out[:200])
    # the source page prints one table per launch; keep it simple: dump the top lines by samples of launch `kid`
    tables, cur = [], []
    for line in out.splitlines():
        if line.startswith('"#"') or line.startswith('"Address"') or line.startswith('"Source"'):
            if cur:
                tables.append(cur)
            cur = [line]
        elif cur:
            cur.append(line)
    if cur:
        tables.append(cur)
    t = tables[int(kid)]
    rows = list(csv.reader(io.StringIO("\n".join(t))))
    hdr = rows[0]
    si = [i for i, h in enumerate(hdr) if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)" or h.startswith("Warp Stall Sampling (All")]
    print(hdr)
    if not si:
        return
    s = si[0]
    body = [r for r in rows[1:] if len(r) > s and r[s].replace(",", "").isdigit()]
    body.sort(key=lambda r: -int(r[s].replace(",", "")))
    tot = sum(int(r[s].replace(",", "")) for r in body) or 1
    for r in body[:int(top)]:
        print(f"{int(r[s].replace(',', '')) / tot * 100:5.1f}%  {r[0][:14]:14s} {r[1][:110]}")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "full":
        full(sys.argv[2])
    else:
        source(*sys.argv[2:])
