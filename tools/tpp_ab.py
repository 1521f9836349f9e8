#!/usr/bin/env python
"""A/B of the two solve paths on one synthetic batch: thread-per-problem (large batches) vs warp-per-problem.
    python tools/tpp_ab.py [B] [N] [reps] [start]      start: zero (default) | warm (the problem's own solution) | rollout
Prints kernel ms of each path and how their results differ (status, iterations, first input)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
start = sys.argv[4] if len(sys.argv) > 4 else "zero"
dev = torch.device("cuda", 0)
b = workload.make_batch(B, N)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
out = {}
for name, thr in (("warp", "0"), ("tpp", "1")):
    os.environ["MPCB200_TPP_MIN_BATCH"] = thr
    s = capi.Solver(N, start_mode=capi.START_ROLLOUT if start == "rollout" else capi.START_ZERO)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    s.set_stream(st.cuda_stream)
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
    cost = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    traj = torch.empty((B, 6 * N + 4), dtype=torch.float64, device=dev)
    best = 1e9
    sol = None
    if start == "warm":
        s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], status=status, traj=traj)
        torch.cuda.synchronize()
        sol = traj.clone()
    for r in range(reps):
        warm = sol.clone() if sol is not None else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], warm=warm, cost=cost, status=status, iters=iters, traj=traj)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = dict(u0=u0.cpu().numpy(), cost=cost.cpu().numpy(), status=status.cpu().numpy(), iters=iters.cpu().numpy(), traj=traj.cpu().numpy(), ms=best)
    it = out[name]["iters"]; stt = out[name]["status"]
    print("%-5s %.3f ms  B=%d N=%d  conv=%.5f  mean_iters=%.2f  solves/s=%.0f" % (name, best, B, N, (stt == 0).mean(), it.mean(), (stt == 0).sum() / best * 1e3), flush=True)
w, t = out["warp"], out["tpp"]
both = (w["status"] == 0) & (t["status"] == 0)
print("status equal %.6f  iters equal %.6f  max|du| (both Optimal) %.3e  >1e-5: %d  max rel cost %.3e  traj max %.3e" % (
    (w["status"] == t["status"]).mean(), (w["iters"] == t["iters"]).mean(), np.abs(w["u0"] - t["u0"])[both].max(),
    (np.abs(w["u0"] - t["u0"])[both].max(axis=1) > 1e-5).sum(),
    (np.abs(w["cost"] - t["cost"]) / np.maximum(1, np.abs(w["cost"])))[both].max(), np.abs(w["traj"] - t["traj"])[both].max()))
print("speed-up %.2fx" % (w["ms"] / t["ms"]))
