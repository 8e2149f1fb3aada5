#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_<tag>.csv  > profiles/<round>_launches_<tag>.txt
    python tools/ncu_summary.py full     gpurun_out/prof_<tag>.ncu-rep  > profiles/<round>_ncu_<tag>.txt
    python tools/ncu_summary.py source   gpurun_out/prof_<tag>.ncu-rep <kernel-id> [top]   (hot SASS/source lines)
    python tools/ncu_summary.py blocks   gpurun_out/prof_<tag>.ncu-rep <kernel name>        (issue share / active lanes per basic block)
    python tools/ncu_summary.py lines    gpurun_out/prof_<tag>.ncu-rep <launch index> [top]  (executed warp instructions per CUDA source line)
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
        if name.startswith("void "):          # templated kernels: "void <unnamed>::k_surface<(bool)0>" -> "<unnamed>::k_surface<0>", "void cub::X<...>" -> "void cub::X"
            body = name[5:]
            m = re.match(r"(<unnamed>::\w+)<\(?(?:bool\)?)?(\w+)>$", body)
            name = f"{m.group(1)}<{m.group(2)}>" if m else "void " + re.sub(r"<.*", "", body)
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


def lines(path, launch, top=40):
    """executed warp instructions, active lanes and stall samples per CUDA source line of one captured launch (needs -lineinfo + --import-source)"""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    cur, fn, agg, tw, tt = None, None, [], 0, 0
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            fn = r[1]
        elif r[0] not in ("Line No", ""):
            try:
                w, t, smp = int(r[7]), int(r[8]), int(r[6])
            except (ValueError, IndexError):
                continue
            agg.append((w, t, smp, cur, r[0], r[1].strip()[:100]))
            tw += w
            tt += t
    print(f"# {path} launch {launch}: {fn}")
    print(f"# warp instructions {tw}, thread instructions {tt}, active lanes {tt / max(tw, 1):.1f}")
    print("# share  cumulative  lanes  stall-samples  file:line  source")
    cum = 0
    for w, t, smp, f, ln, src in sorted(agg, reverse=True)[:int(top)]:
        cum += w
        print(f"{100 * w / tw:5.1f}%  {100 * cum / tw:5.1f}%  {t / max(w, 1):5.1f}  {smp:5d}  {f}:{ln}  {src}")


def blocks(path, kernel, min_share=0.004):
    """Basic-block profile of one kernel from the source page of an `--import-source on` capture: for every run of SASS
    instructions with the same execution count, its share of all issued warp instructions, the average number of active
    lanes, its share of the stall samples and its opcode mix. This is what tells which PHASE of the persistent traversal loop
    (fetch / node step / triangle test / pop) the issue slots go to and how full its warps are."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-kernel-base", "function"], capture_output=True, text=True).stdout
    tables, cur = [], None
    for ln in out.splitlines():
        if ln.startswith('"Kernel Name"'):
            cur = {"name": ln, "rows": []}
            tables.append(cur)
        elif cur is not None:
            cur["rows"].append(ln)
    for t in tables:
        if kernel not in t["name"]:
            continue
        rows = list(csv.reader(io.StringIO("\n".join(t["rows"]))))
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}
        body = [r for r in rows[1:] if len(r) == len(hdr)]

        def num(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.0
        tot_i = sum(num(r[ix["Instructions Executed"]]) for r in body)
        tot_t = sum(num(r[ix["Thread Instructions Executed"]]) for r in body)
        print(f"# {path}: {t['name']}  ({len(body)} SASS instructions)")
        print(f"# warp instructions {tot_i:.4g}, thread instructions {tot_t:.4g}, average active lanes {tot_t / max(tot_i, 1):.2f}")
        blks = []
        for r in body:
            ie, te, sm = num(r[ix["Instructions Executed"]]), num(r[ix["Thread Instructions Executed"]]), num(r[ix["# Samples"]])
            op = r[1].split()[0] if r[1].split() else ""
            if blks and blks[-1]["ie"] == ie and not r[1].strip().startswith(("BRA", "@")):
                b = blks[-1]; b["n"] += 1; b["te"] += te; b["sm"] += sm; b["last"] = r[0]; b["ops"].append(op)
            else:
                blks.append({"first": r[0], "last": r[0], "ie": ie, "n": 1, "te": te, "sm": sm, "ops": [op]})
        tot_s = sum(b["sm"] for b in blks) or 1
        print(f"{'addresses':13s} {'instr':>5s} {'executed':>10s} {'issue share':>11s} {'lanes':>6s} {'samples':>8s}  opcode mix")
        for b in blks:
            w = b["ie"] * b["n"]
            if w / max(tot_i, 1) < min_share:
                continue
            mix = collections.Counter(o.split(".")[0] for o in b["ops"]).most_common(6)
            print(f"{b['first'][-5:]}-{b['last'][-5:]:7s} {b['n']:5d} {b['ie']:10.4g} {100 * w / tot_i:10.1f}% {b['te'] / max(w, 1):6.1f} {100 * b['sm'] / tot_s:7.1f}%  "
                  + " ".join(f"{k}:{v}" for k, v in mix))
        return


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "full":
        full(sys.argv[2])
    elif mode == "blocks":
        blocks(sys.argv[2], sys.argv[3])
    elif mode == "lines":
        lines(*sys.argv[2:])
    else:
        source(*sys.argv[2:])
