#!/usr/bin/env python
"""Small fixed workload for ncu: `python tools/prof_run.py [B] [N] [reps] [warm]` solves one synthetic
batch `reps` times through the C ABI (device pointers) and prints kernel ms per launch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
warm_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
FRENET = os.environ.get("PROF_MODEL") == "frenet"   # the Frenet-frame variant (mpc_solve_frenet_kernel)
b = workload.make_frenet_batch(B, N) if FRENET else workload.make_batch(B, N)
if FRENET:
    b["ref"] = b["kpoly"]
if os.environ.get("PROF_SAME"):   # every problem = problem j: all warps run the same instruction stream
    j = int(os.environ["PROF_SAME"])
    for k in ("state", "ref", "u_prev", "v_des"):
        b[k] = np.ascontiguousarray(np.repeat(b[k][j:j + 1], B, axis=0))
s = capi.FrenetSolver(N) if FRENET else capi.Solver(N)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
s.set_stream(st.cuda_stream)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
cost = torch.empty(B, dtype=torch.float64, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev)
iters = torch.empty(B, dtype=torch.int32, device=dev)
warm = None
if warm_mode:
    traj = torch.zeros((B, 6 * N + 4), dtype=torch.float64, device=dev)
    s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], status=status, traj=traj)
    torch.cuda.synchronize()
for r in range(reps):
    if warm_mode:
        warm = traj.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], warm=warm, cost=cost, status=status, iters=iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = iters.cpu().numpy(); stt = status.cpu().numpy()
    print("rep %d: %.3f ms  B=%d N=%d  conv=%.4f  mean_iters=%.2f  solves/s=%.0f  iters/s=%.3e" %
          (r, ms, B, N, (stt == 0).mean(), it.mean(), (stt == 0).sum() / ms * 1e3, it.sum() / ms * 1e3))
