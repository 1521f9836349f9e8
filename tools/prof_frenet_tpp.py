#!/usr/bin/env python
"""Small fixed workloads of the thread-per-problem kernels for ncu:
    python tools/prof_frenet_tpp.py batch [B] [N]        mpc_solve_tpp_kernel<1> on a Frenet batch
    python tools/prof_frenet_tpp.py rollout [V] [T]      the XY closed-loop pipeline (plant / waypoints / solve per period)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "batch"
if mode == "batch":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    b = workload.make_frenet_batch(B, N)
    s = capi.FrenetSolver(N)
    s.set_large_batch_path(1)
    for r in range(2):
        g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"])
        print("rep %d: kernel %.3f ms, converged %.4f, mean iters %.2f" % (r, s.stats()["kernel_ms"], (g["status"] == 0).mean(), g["iters"].mean()))
else:
    from rollout_bench import fleet
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    trajs, path_of, pose0 = fleet(V)
    s = capi.Solver(8)
    s.set_large_batch_path(1)
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
    out = s.rollout(pose0, path_of, T)
    print("kernels %.3f ms for %d periods, optimal %.4f, mean iters %.2f" % (s.stats()["kernel_ms"], T, (out["log"][:, :, 6] == 0).mean(), out["log"][:, :, 7].mean()))
