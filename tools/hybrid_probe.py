#!/usr/bin/env python
"""Experiment: the two kernel layouts side by side on one GPU.  The warp-per-problem kernel (on-chip, issue/MIO bound, no
DRAM traffic) with 2 blocks per SM and the thread-per-problem kernel (HBM-streaming, few issue slots) with one 64-thread
block per SM fit on an SM together (registers: 2 x 128 x 168 + 64 x 255 = 59 K).  A batch is split statically: the first
`frac` of the problems go to the thread-per-problem kernel.
    python tools/hybrid_probe.py [B] [N] [frac ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
fracs = [float(x) for x in sys.argv[3:]] or [0.0, 0.2, 0.3, 0.4]
dev = torch.device("cuda", 0)
b = workload.make_batch(B, N)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
os.environ["MPCB200_BLOCKS_PER_SM"] = os.environ.get("HYB_WARP_BLOCKS", "2")
sw = capi.Solver(N); sw.set_large_batch_path(0)
os.environ["MPCB200_TPP_BLOCK"] = os.environ.get("HYB_TPP_BLOCK", "64")
stp = capi.Solver(N); stp.set_large_batch_path(1)
del os.environ["MPCB200_BLOCKS_PER_SM"]
s3 = capi.Solver(N); s3.set_large_batch_path(0)     # the plain configuration, for reference
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
sw.set_stream(s1.cuda_stream); stp.set_stream(s2.cuda_stream); s3.set_stream(s1.cuda_stream)
u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev)
iters = torch.empty(B, dtype=torch.int32, device=dev)


def run(frac, plain=False):
    nt = int(B * frac) // 256 * 256
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    e0.record(s1)
    s2.wait_event(e0)
    if plain:
        s3.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], status=status, iters=iters)
    else:
        if B - nt:
            sw.solve_batch_device(B - nt, d["state"][nt:], d["ref"][nt:], d["u_prev"][nt:], u0[nt:], v_des=d["v_des"][nt:], status=status[nt:], iters=iters[nt:])
        if nt:
            stp.solve_batch_device(nt, d["state"][:nt], d["ref"][:nt], d["u_prev"][:nt], u0[:nt], v_des=d["v_des"][:nt], status=status[:nt], iters=iters[:nt])
    e1.record(s1); e2.record(s2)
    s1.wait_event(e2)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), e0.elapsed_time(e2)


run(0.0, plain=True)
t = min(run(0.0, plain=True)[0] for _ in range(2))
print("plain warp-per-problem (3 blocks/SM): %.2f ms  %.0f solves/s" % (t, B / t * 1e3), flush=True)
for f in fracs:
    best = None
    for _ in range(2):
        tw, tt = run(f)
        if best is None or max(tw, tt) < max(best): best = (tw, tt)
    ok = (status.cpu().numpy() == 0).mean()
    print("frac %.2f to thread-per-problem: warp part %.2f ms, thread part %.2f ms -> %.2f ms  %.0f solves/s  conv %.5f" %
          (f, best[0], best[1], max(best), B / max(best) * 1e3, ok), flush=True)
