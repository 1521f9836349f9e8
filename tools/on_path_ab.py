#!/usr/bin/env python
"""mpcb200_solve_batch_on_path (host buffers) at N = 8: thread-per-problem (default rule) vs warp-per-problem kernel.
    python tools/on_path_ab.py [B] [N]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402
from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
b = workload.make_batch(B, N)
path_of = (b["path"] - 1).astype(np.int32)
trajs = [GPSRefTrajectory(mat_filename=p, traj_horizon=N, traj_dt=0.2) for p in (1, 2, 3)]
res = {}
for name, mb in (("warp", 0), ("thread", -1)):
    s = capi.Solver(N)
    s.set_large_batch_path(mb)
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
    best = (1e9, 1e9)
    for r in range(3):
        t = time.perf_counter()
        g = s.solve_batch_on_path(b["state"], path_of, b["u_prev"], v_des=b["v_des"])
        wall = time.perf_counter() - t
        best = min(best, (s.stats()["kernel_ms"], wall * 1e3))
    res[name] = g
    print("%-6s kernels %.3f ms (%d launches), call %.3f ms, optimal %.5f, %.0f solves/s end to end" % (
        name, best[0], s.stats()["kernel_launches"], best[1], (g["status"] == 0).mean(), (g["status"] == 0).sum() / best[1] * 1e3), flush=True)
w, t = res["warp"], res["thread"]
both = (w["status"] == 0) & (t["status"] == 0)
print("status equal %.6f, stop equal %s, max|du| %.2e" % ((w["status"] == t["status"]).mean(), np.array_equal(w["stop"], t["stop"]), np.abs(w["u0"] - t["u0"])[both].max()))
