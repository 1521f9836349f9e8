#!/usr/bin/env python
"""Where a single solve's latency goes: python tools/latency_probe.py [N]
wall time of capi.Solver.solve_batch for a batch of one (H2D + kernel + D2H, host buffers) vs the device time
of the kernel alone (CUDA events inside the library)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
b = workload.make_batch(64, N)
s = capi.Solver(N)
g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
for mode in ("warm", "cold"):
    wall, kern, its = [], [], []
    for i in range(300):
        j = i % 64
        w = g["traj"][j:j + 1].copy() if mode == "warm" else None
        t0 = time.perf_counter()
        r = s.solve_batch(b["state"][j:j + 1], b["ref"][j:j + 1], b["u_prev"][j:j + 1], v_des=b["v_des"][j:j + 1], warm=w)
        wall.append(time.perf_counter() - t0)
        kern.append(s.stats()["kernel_ms"]); its.append(int(r["iters"][0]))
    wall, kern = 1e3 * np.array(wall[50:]), np.array(kern[50:])
    print("N=%d %s: wall p50 %.3f ms p99 %.3f | kernel p50 %.3f ms | overhead p50 %.3f ms | iters mean %.1f" %
          (N, mode, np.percentile(wall, 50), np.percentile(wall, 99), np.percentile(kern, 50), np.percentile(wall - kern, 50), np.mean(its[50:])))
