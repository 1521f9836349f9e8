#!/usr/bin/env python
"""Horizon sweep (BASELINE.json configs[4]): `python tools/horizon_sweep.py [B] [start]` solves B synthetic
problems at N = 8, 20, 40, 80 through the C ABI (device pointers) and prints solves/s.
start = zero | ref | rollout  (all-zero `start=0.0`, the reference waypoints as start point, or
MPCB200_START_ROLLOUT: previous command rolled out from the measured state)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
start = sys.argv[2] if len(sys.argv) > 2 else "ref"
dev = torch.device("cuda", 0)
for N in (8, 20, 40, 80):
    b = workload.make_batch(B, N)
    s = capi.Solver(N, start_mode=capi.START_ROLLOUT if start == "rollout" else capi.START_ZERO)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    s.set_stream(st.cuda_stream)
    d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
    w0 = torch.from_numpy(workload.reference_start(b, N)).to(dev) if start == "ref" else None
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    best = 1e30
    for r in range(3):
        warm = w0.clone() if w0 is not None else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], warm=warm, status=status, iters=iters)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    stt = status.cpu().numpy(); it = iters.cpu().numpy()
    print("N=%d B=%d start=%s: %.3f ms  conv=%.4f  mean_iters=%.2f  converged solves/s=%.0f  iters/s=%.3e" %
          (N, B, start, best, (stt == 0).mean(), it.mean(), (stt == 0).sum() / best * 1e3, it.sum() / best * 1e3))
