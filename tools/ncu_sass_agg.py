#!/usr/bin/env python
"""Aggregate the SASS-level source page of an .ncu-rep (captured with --import-source on) per opcode:
   python tools/ncu_sass_agg.py report.ncu-rep <kernel regex> [top-N instructions]"""
import csv, collections, re, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not hi:
    sys.exit("no SASS page")
# first launch only
start = hi[0]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return int(float(r[ix[k]]))
    except Exception: return 0
tot = collections.Counter(); wf = collections.Counter(); wfi = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
per = []
for r in rows[start + 1:end]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9]+)?)', src)
    op = m.group(1) if m else '?'
    ie = num(r, 'Instructions Executed')
    tot[op] += ie; thr[op] += num(r, 'Thread Instructions Executed')
    w, wi = num(r, 'L1 Wavefronts Shared'), num(r, 'L1 Wavefronts Shared Ideal')
    wf[op] += w; wfi[op] += wi
    s = num(r, '# Samples'); samp[op] += s
    per.append((s, ie, w, wi, src[:90]))
n = sum(tot.values()); ns = max(1, sum(samp.values()))
print("warp instructions %d, thread instructions %d, stall samples %d" % (n, sum(thr.values()), ns))
for k, v in tot.most_common(30):
    print("%-12s inst %5.1f%%  samples %5.1f%%  smem wavefronts %d (ideal %d)" % (k, 100.0 * v / n, 100.0 * samp[k] / ns, wf[k], wfi[k]))
print("shared wavefronts total %d ideal %d" % (sum(wf.values()), sum(wfi.values())))
per.sort(reverse=True)
print("-- top instructions by stall samples: samples, executed, wavefronts, ideal, sass")
for p in per[:top]: print(p)
