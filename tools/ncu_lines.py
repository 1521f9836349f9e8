#!/usr/bin/env python
"""Aggregate an ncu report's SASS page per CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [libmpc_b200.so] [top]

ncu's CSV export carries metrics only for the SASS view; nvdisasm -g gives the line of every
instruction of the same cubin, in the same (address) order, so the two are zipped by offset."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "mkz_mpc_path_follower_b200", "libmpc_b200.so")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
kernel = sys.argv[4] if len(sys.argv) > 4 else "mpc_solve_kernel"

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# offset -> (file, line) for the kernel's section
line_of = {}
cur = None
insec = False
for ln in dis.splitlines():
    if ln.startswith("//----") and ".text." in ln:
        insec = kernel in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if "inlined at" not in ln or cur is None:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        else:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
cw = hdr.index("L1 Wavefronts Shared Excessive")
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_i = tot_s = 0
opc = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key, sass = line_of.get(addr - base, (("?", 0), r[1]))
    n, s = int(r[ci] or 0), int(r[cs] or 0)
    a = agg[key]
    a[0] += n; a[1] += s; a[2] += int(r[cw] or 0)
    for h, i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[3][h] += v
    tot_i += n; tot_s += s
    opc[r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]] += n
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("top opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in opc.most_common(14)))
src_cache = {}


def src(f, l):
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mkz_mpc_path_follower_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""


for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join("%s:%d" % (k.replace("stall_", ""), v) for k, v in a[3].most_common(3))
    print("%5.2f%% inst %5.2f%% smp excwf=%-10d %s:%d  [%s]  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / max(1, tot_s), a[2], key[0], key[1], st, src(*key)))

# ---- region table (line ranges of mpc_kernel.cuh), normalised per solver iteration if given
if os.environ.get("NCU_REGIONS"):
    regs = []
    for part in os.environ["NCU_REGIONS"].split(","):
        name, a, b = part.split(":")
        regs.append((name, int(a), int(b)))
    per = float(os.environ.get("NCU_PER", "1"))
    acc = collections.Counter(); accs = collections.Counter()
    for (f, l), a in agg.items():
        nm = "other:" + f
        if f == "mpc_kernel.cuh":
            for name, lo, hi in regs:
                if lo <= l <= hi:
                    nm = name
                    break
        acc[nm] += a[0]; accs[nm] += a[1]
    print("\nregion table (instructions per unit = total / %g):" % per)
    for nm, v in acc.most_common():
        print("  %-28s %6.2f%%  %9.1f inst/unit   %5.2f%% samples" % (nm, 100.0 * v / tot_i, v / per, 100.0 * accs[nm] / max(1, tot_s)))
