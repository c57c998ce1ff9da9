#!/usr/bin/env python
"""Table of the bench lines kept under profiles/ (one JSON line per file):  python profiles/collect.py r02 > table.md"""
import glob, json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
here = os.path.dirname(os.path.abspath(__file__))
rows = []
for f in sorted(glob.glob(os.path.join(here, tag + "_bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    if d.get("impl") == "reference":
        rows.append((os.path.basename(f), "CPU `advance()` port, %d threads" % d["cpu_baseline"]["cores"], "-", "%.1f M" % (d["value"] / 1e6),
                     "%.1f" % d["ms_per_step"], "-", "-", "-", "single thread %.1f M" % (d["cpu_baseline"]["single_thread_value"] / 1e6)))
        continue
    c, r = d["config"], d["roofline"]
    ws = r["whole_substep"]
    note = "kernel %s %.3f ms (%.0f %% of measured HBM)" % (r["kernel"], r["kernel_ms"], 100 * r["frac"])
    if d["n_gpus"] > 1:
        note += "; bubble %.2f ms; %s" % (d.get("engine", c).get("exchange_bubble_ms_per_step", 0), d["scaling"])
    rows.append((os.path.basename(f), "%s, %.1f M particles, n_grid %d" % (c["name"], c["particles"] / 1e6, c["n_grid"]), d["n_gpus"],
                 "%.2f G" % (d["value"] / 1e9), "%.3f" % d["ms_per_step"], "%.0f" % ws["achieved"], "%.1f %%" % (100 * ws["frac"]),
                 "%.1f %%" % (100 * ws["frac_of_nominal_8TBs"]), note + "; e2e %.2f G" % (d["e2e"]["value"] / 1e9)))
print("| file | workload | GPUs | particle-substeps/s | ms/substep | algorithmic GB/s per GPU | of measured 6.55 TB/s | of nominal 8 TB/s | notes |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(str(x) for x in r) + " |")
