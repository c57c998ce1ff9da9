#!/usr/bin/env python
"""Static SASS statistics of one kernel of libmpm.so: opcode histogram and instructions per source line.

    python tools/sass_lines.py <substring of the mangled kernel name> [--so path] [--top N]

Uses cuobjdump -xelf + nvdisasm -g (needs -lineinfo at compile time, which csrc/Makefile passes).
Static counts only: loops are counted once -- read them next to the kernel's structure.
"""
import argparse, collections, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpm_flip98a_b200", "libmpm.so"))
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--cubin", default="mpm_kernels")
a = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith(a.cubin)][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur_fn, cur_line, fn_hit = None, None, False
ops, lines = collections.Counter(), collections.Counter()
total = 0
for ln in out.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        cur_fn = m.group(1)
        fn_hit = a.kernel in cur_fn
        continue
    if not fn_hit:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", ln)
    if m:
        op = m.group(1)
        ops[op.split(".")[0]] += 1
        lines[cur_line] += 1
        total += 1
print("kernel match:", a.kernel, "static instructions:", total)
print("-- opcodes")
for k, v in ops.most_common(a.top):
    print("%6d  %s" % (v, k))
print("-- source lines")
for k, v in lines.most_common(a.top):
    print("%6d  %s:%s" % (v, k[0] if k else "?", k[1] if k else "?"))
