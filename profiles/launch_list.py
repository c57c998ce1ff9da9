#!/usr/bin/env python
"""ncu launch list (csv of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv <cmd>`) ->
markdown table of per-kernel launches / total / mean / share.   python profiles/launch_list.py x.csv [first] [last] > out.md"""
import collections, csv, re, sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])
    name = re.sub(r"\bmpm::|\(anonymous namespace\)::|void |<unnamed>::", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    rows.append((name, us))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[first:last]
agg = collections.OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
for n, (c, t) in agg.items():
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (n, c, t, t / c, 100 * t / tot))
print("| **sum** | %d | %.1f | | 100%% |" % (len(rows), tot))
