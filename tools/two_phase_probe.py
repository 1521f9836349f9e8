#!/usr/bin/env python
"""Probe of a two-phase cold-start solve: the thread-per-problem kernel with an iteration cap K on every problem, then the
warp-per-problem kernel from scratch on the problems that hit the cap.
    python tools/two_phase_probe.py [B] [N] [K ...]
Prints the kernel ms of each phase and of the warp kernel alone on the whole batch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
Ks = [int(a) for a in sys.argv[3:]] or [50, 60, 70, 80]
dev = torch.device("cuda", 0)
b = workload.make_batch(B, N)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}


def run(s, idx=None, reps=3):
    n = B if idx is None else int(idx.numel())
    inp = d if idx is None else {k: v[idx].contiguous() for k, v in d.items()}
    u0 = torch.empty((n, 2), dtype=torch.float64, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    iters = torch.empty(n, dtype=torch.int32, device=dev)
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.solve_batch_device(n, inp["state"], inp["ref"], inp["u_prev"], u0, v_des=inp["v_des"], status=status, iters=iters)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, u0, status, iters


st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
warp = capi.Solver(N)
warp.set_stream(st.cuda_stream)
warp.set_large_batch_path(0)
w_ms, w_u0, w_status, w_iters = run(warp)
print("warp alone: %.3f ms, mean iters %.2f, optimal %.5f" % (w_ms, w_iters.float().mean().item(), (w_status == 0).float().mean().item()), flush=True)
for K in Ks:
    tp = capi.Solver(N, max_iter=K)
    tp.set_stream(st.cuda_stream)
    tp.set_large_batch_path(1)
    t_ms, t_u0, t_status, t_iters = run(tp)
    idx = torch.nonzero(t_status != 0).flatten()
    r_ms, r_u0, r_status, r_iters = run(warp, idx) if idx.numel() else (0.0, None, None, None)
    u0 = t_u0.clone()
    status = t_status.clone()
    if idx.numel():
        u0[idx] = r_u0
        status[idx] = r_status
    both = (status == 0) & (w_status == 0)
    print("K=%3d: thread-per-problem %.3f ms + warp on %d stragglers %.3f ms = %.3f ms (%.2fx);  status equal %.6f, max|du| %.2e" % (
        K, t_ms, idx.numel(), r_ms, t_ms + r_ms, w_ms / (t_ms + r_ms), (status == w_status).float().mean().item(),
        (u0 - w_u0).abs()[both].max().item()), flush=True)
    tp.close()
