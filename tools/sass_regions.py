#!/usr/bin/env python
"""Static (and, with an ncu report, dynamic) instruction counts of a kernel per solver region.

    python tools/sass_regions.py [--kernel mpc_solve_kernel] [--rep gpurun_out/prof.ncu-rep] [--per ITERS]

Every SASS instruction is attributed through its inline chain (nvdisasm -gi) to the innermost
solver routine of mpc_kernel.cuh that contains it (eval_point, assemble, Riccati passes, ...) or,
for code inlined straight into solve(), to the driver phase.  Regions are found from markers in
the source, so the table follows the file as it changes."""
import argparse
import collections
import csv
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "mkz_mpc_path_follower_b200", "csrc", "mpc_kernel.cuh")
ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="mpc_solve_kernel")
ap.add_argument("--so", default=os.path.join(ROOT, "mkz_mpc_path_follower_b200", "libmpc_b200.so"))
ap.add_argument("--rep", default=None)
ap.add_argument("--per", type=float, default=1.0)
args = ap.parse_args()

src = open(SRC).read().splitlines()


def find(pat, start=0):
    for i in range(start, len(src)):
        if re.search(pat, src[i]):
            return i + 1
    raise SystemExit("marker not found: " + pat)


marks = [  # (region, first line); a region ends where the next one starts
    ("kappa_clamp", find(r"MPC_DEV_NOINLINE void kappa_sigma_clamp")),
    ("roles_fn", find(r"MPC_HD void riccati_roles")),
    ("team_collectives", find(r"MPC_DEV void tsync\(\)")),
    ("init_work", find(r"static void init_work")),
    ("recips", find(r"MPC_DEV void recips")),
    ("eval_point", find(r"MPC_DEV void eval_point")),
    ("gradient", find(r"struct Grad ")),
    ("assemble", find(r"MPC_DEV void assemble")),
    ("backward_wrap", find(r"MPC_DEV bool riccati_backward\(\)")),
    ("backward_roles", find(r"MPC_DEV bool riccati_backward_warp")),
    ("backward_loop", find(r"for \(int s = N - 1; s >= 0; s--\)")),
    ("forward", find(r"MPC_DEV void riccati_forward")),
    ("duals", find(r"MPC_DEV void recover_duals")),
    ("alpha_primal", find(r"MPC_DEV double alpha_primal")),
    ("rollout_restore", find(r"MPC_DEV void rollout_restore")),
    ("nlp_feasible", find(r"MPC_DEV bool nlp_feasible")),
    ("solve_init", find(r"MPC_DEV Result solve\(\)")),
    ("drv_heavy_dispatch", find(r"shared heavy work")),
    ("drv_trial_test", find(r"if \(phase == PH_TRIAL \|\| phase == PH_SOC\)")),
    ("drv_accept", find(r"---------------- accepted")),
    ("drv_resto_init", find(r"if \(phase == PH_RESTO\)")),
    ("drv_eval0_ls", find(r"if \(phase == PH_EVAL0\)")),
    ("drv_begin", find(r"if \(phase == PH_BEGIN\)")),
    ("drv_resolve_pd", find(r"if \(phase == PH_RESOLVE_EVAL\)")),
    ("solve_exit", find(r"res.iters = iter;")),
    ("io", find(r"Problem I/O for one warp")),
    ("rollout", find(r"Closed-loop rollout")),
]
LEAF = {"kappa_clamp", "rollout_restore", "recips", "eval_point", "gradient", "assemble", "backward_roles", "backward_loop", "backward_wrap", "forward", "duals",
        "alpha_primal", "nlp_feasible", "init_work"}


def region_of(line):
    r = "pre"
    for name, l0 in marks:
        if line >= l0:
            r = name
    return r


tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(args.so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
insec = False
chain = []
fresh = True
reg_of_off = {}
for ln in dis.splitlines():
    if ln.startswith("//----") and ".text." in ln:
        insec = args.kernel in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        fresh = True
        regs = [region_of(l) for f, l in chain if f == "mpc_kernel.cuh"]
        leaf = [r for r in regs if r in LEAF]
        if leaf:
            # innermost leaf routine, except that helpers (recips, gradient) count for their caller
            pick = leaf[0]
            for r in leaf:
                if r not in ("recips", "gradient"):
                    pick = r
                    break
        else:
            drv = [r for r in regs if r not in ("team_collectives", "pre")]
            pick = drv[-1] if drv else (regs[-1] if regs else "other")
            if pick in ("io", "rollout"):
                pick = drv[0]
        reg_of_off[int(m.group(1), 16)] = pick

static = collections.Counter(reg_of_off.values())
dyn = collections.Counter()
smp = collections.Counter()
if args.rep:
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
    base = None
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        reg = reg_of_off.get(addr - base, "other")
        dyn[reg] += int(r[ci] or 0); smp[reg] += int(r[cs] or 0)
tot = sum(static.values())
print("kernel %s: %d SASS instructions (%.1f KB)" % (args.kernel, tot, tot * 16 / 1024.0))
td, ts = max(1, sum(dyn.values())), max(1, sum(smp.values()))
print("%-22s %8s %7s %12s %7s %8s" % ("region", "static", "KB", "dyn/unit", "dyn%", "samples%"))
for name, n in sorted(static.items(), key=lambda kv: -(dyn[kv[0]] if args.rep else kv[1])):
    print("%-22s %8d %7.1f %12.1f %6.2f%% %7.2f%%" % (name, n, n * 16 / 1024.0, dyn[name] / args.per, 100.0 * dyn[name] / td, 100.0 * smp[name] / ts))
