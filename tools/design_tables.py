#!/usr/bin/env python
"""Prints the measurement tables of DESIGN.md §5 / §6 from the committed bench lines under profiles/ (so that the document quotes the files)."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    rows = [l for l in open(p).read().splitlines() if l.startswith("{")]
    return json.loads(rows[-1]) if rows else None


def main():
    """`--write`: replaces the blocks between <!-- TABLE:x --> and <!-- /TABLE:x --> in DESIGN.md (x = main, scaling, image)"""
    import contextlib
    import io
    import re
    write = "--write" in sys.argv
    args = [a for a in sys.argv[1:] if a != "--write"]
    prefix = args[0] if args else "r02_final_"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tables(prefix)
    parts = buf.getvalue().strip("\n").split("\n\n")
    if not write:
        print(buf.getvalue())
        return
    p = os.path.join(ROOT, "DESIGN.md")
    doc = open(p).read()
    for name, body in zip(("main", "scaling", "image"), parts):
        doc = re.sub(rf"<!-- TABLE:{name} -->.*?<!-- /TABLE:{name} -->", lambda m: f"<!-- TABLE:{name} -->\n{body}\n<!-- /TABLE:{name} -->", doc, flags=re.S)
    open(p, "w").write(doc)


def tables(prefix):
    print("| workload | Mpaths/s | e2e | Mrays/s | `roofline.frac` | extend Grays/s | CPU (16 cores) | file |")
    print("|---|---|---|---|---|---|---|---|")
    for wl, f in (("C1", "bench_c1"), ("C2", "bench_c2"), ("**C3** (headline)", "bench_c3"), ("C4 `ptdirect` 256 spp", "bench_c4"), ("C4 `pt` 256 spp", "bench_c4pt"),
                  ("C2's scene `bdpt`", "bench_c2bdpt")):
        j = line(prefix + f + ".json")
        if not j:
            continue
        r = j.get("roofline") or {}
        c = j.get("cpu_baseline") or {}
        print(f"| {wl} | {j['value']:.0f} | {j['e2e']['value']:.0f} | {j['mrays_per_s']:.0f} | {r.get('frac', 0):.2f} | {r.get('grays_per_s', 0):.2f} | "
              f"{(str(round(c['value'], 2)) + ' (' + c['kind'] + ')') if c else '—'} | `{prefix + f}.json` |")
    for f in ("bench_ref_c3", "bench_ref_c2"):
        j = line(prefix + f + ".json")
        if j:
            print(f"| `--impl reference` {f[-2:].upper()} | {j['value']:.2f} | | {j['mrays_per_s']:.1f} | | | {j['cpu_baseline']['cores']} cores, {j['cpu_baseline']['kind']} | `{prefix + f}.json` |")
    for f in ("raybench_c3_16M", "raybench_c4_16M"):
        j = line(prefix + f + ".json")
        if j:
            b = j["batches"]
            print(f"| C5 {f} | | | " + " / ".join(f"{b[k]['grays_per_s']:.2f}" for k in ("coherent", "incoherent", "coherent_occlusion", "incoherent_occlusion")) +
                  f" Grays/s | | | mismatches {sum(b[k]['mismatches'] for k in b)} of {sum(b[k]['checked'] for k in b)} | `{prefix + f}.json` |")
    print()
    print("| N | Mpaths/s | ms / step | efficiency | e2e Mpaths/s | e2e efficiency | file |")
    print("|---|---|---|---|---|---|---|")
    base = None
    for n in (1, 2, 4, 8):
        for pat in (f"r02_bench_c3_n{n}_f5.json", f"r02_bench_c3_n{n}_n8v.json", f"r02_bench_c3_n{n}_n8.json", f"r02_bench_c3_n{n}_s44.json"):
            j = line(pat)
            if j:
                if n == 1:
                    base = j
                eff = j["value"] / (n * base["value"]) if base else float("nan")
                eff2 = j["e2e"]["value"] / (n * base["e2e"]["value"]) if base else float("nan")
                print(f"| {n} | {j['value']:.0f} | {j['ms_per_step']:.1f} | {eff:.3f} | {j['e2e']['value']:.0f} | {eff2:.3f} | `{pat}` |")
                break
    print()
    print("| case | relRMSE GPU / CPU (clamped) | difference (s.e.) | unclamped difference (s.e.) | clamped mean GPU − CPU (s.e.) | blocks beyond 3 σ, max z |")
    print("|---|---|---|---|---|---|")
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_image_acceptance_*.json"))):
        j = json.load(open(p))
        c, u, m, b = j["rel_rmse_clamped"], j["rel_rmse_unclamped"], j["image_mean"], j["blocks"]
        print(f"| {j['case']} ({j['width']}×{j['height']}, {j['spp']} spp, {j['renders_per_side']} renders per side) | {c['gpu']:.4f} / {c['cpu']:.4f} | {c['diff_pct_of_cpu']:+.2f} % ({c['standard_error_pct']:.2f} %) | "
              f"{u['diff_pct_of_cpu']:+.1f} % ({u['standard_error_pct']:.1f} %) | {m.get('clamped_gpu_vs_cpu_pct', float('nan')):+.2f} % ({m.get('clamped_pair_standard_error_pct', float('nan')):.2f} %) | "
              f"{b['beyond_3_sigma']} of {b['count']}, {b['z_max']:.2f} |")


if __name__ == "__main__":
    main()
