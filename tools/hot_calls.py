#!/usr/bin/env python
"""Executed CALLs (math-library slow paths, out-of-line helpers) and instruction-fetch stall hot spots of a
kernel, from an ncu report and the .so it was taken with:
    python tools/hot_calls.py gpurun_out/prof.ncu-rep gpurun_out/lib.so [kernel]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so = sys.argv[1], sys.argv[2]
kernel = sys.argv[3] if len(sys.argv) > 3 else "mpc_solve_kernelEN"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
insec = False
cur = None
sym = None
lineof, symof = {}, {}
for l in dis.splitlines():
    if l.startswith("//----") and ".text." in l:
        insec = kernel in l
        continue
    if not insec:
        continue
    m = re.match(r"^(\$__internal[^:]*|_Z[^:]*):", l)
    if m:
        sym = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        lineof[int(m.group(1), 16)] = cur
        symof[int(m.group(1), 16)] = sym
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
cn = [i for i, h in enumerate(hdr) if h.startswith("stall_no_inst") and "Not Issued" not in h]
base = None
calls, per_sym, per_sym_s, noinst = [], collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    off = a - base
    n, s = int(r[ci] or 0), int(r[cs] or 0)
    tot += n; tots += s
    per_sym[symof.get(off)] += n; per_sym_s[symof.get(off)] += s
    if "CALL" in r[1] and n:
        calls.append((n, lineof.get(off), r[1][:40]))
    for i in cn:
        noinst[lineof.get(off)] += int(r[i] or 0)
print("executed instructions per function in the kernel's text:")
for k, v in per_sym.most_common():
    print("  %-66s %6.2f%% inst %6.2f%% samples" % (str(k)[:66], 100.0 * v / tot, 100.0 * per_sym_s[k] / max(1, tots)))
print("executed CALL sites:")
for n, ln, ins in sorted(calls, reverse=True)[:15]:
    print("  %9d  %s  %s" % (n, ln, ins))
print("instruction-fetch stall samples per source line (of %d samples):" % tots)
for k, v in noinst.most_common(12):
    print("  %6d  %s" % (v, k))
