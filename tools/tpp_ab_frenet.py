#!/usr/bin/env python
"""Frenet-frame variant: thread-per-problem vs warp-per-problem kernel on one synthetic batch.
    python tools/tpp_ab_frenet.py [B] [N] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
b = workload.make_frenet_batch(B, N)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "kpoly", "u_prev", "v_des")}
out = {}
for name, mb in (("warp", 0), ("tpp", 1)):
    s = capi.FrenetSolver(N)
    s.set_large_batch_path(mb)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    s.set_stream(st.cuda_stream)
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
    cost = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    traj = torch.empty((B, 6 * N + 4), dtype=torch.float64, device=dev)
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.solve_batch_device(B, d["state"], d["kpoly"], d["u_prev"], u0, v_des=d["v_des"], cost=cost, status=status, iters=iters, traj=traj)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = dict(u0=u0.cpu().numpy(), cost=cost.cpu().numpy(), status=status.cpu().numpy(), iters=iters.cpu().numpy(), traj=traj.cpu().numpy(), ms=best)
    it = out[name]["iters"]; stt = out[name]["status"]
    print("%-5s %.3f ms  B=%d N=%d  conv=%.5f  mean_iters=%.2f  solves/s=%.0f" % (name, best, B, N, (stt == 0).mean(), it.mean(), (stt == 0).sum() / best * 1e3), flush=True)
w, t = out["warp"], out["tpp"]
both = (w["status"] == 0) & (t["status"] == 0)
print("status equal %.6f  iters equal %.6f  max|du| (both Optimal) %.3e  traj max %.3e" % (
    (w["status"] == t["status"]).mean(), (w["iters"] == t["iters"]).mean(), np.abs(w["u0"] - t["u0"])[both].max(), np.abs(w["traj"] - t["traj"])[both].max()))
print("speed-up %.2fx" % (w["ms"] / t["ms"]))
