#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel of the library once, small sizes.
    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402
from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory  # noqa: E402

for N, B in ((8, 24), (20, 24), (40, 8), (80, 4)):
    s = capi.Solver(N, max_iter=30)
    b = workload.make_batch(B, N)
    w = workload.reference_start(b, N)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], warm=w, want_traj=True)
    g0 = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    print("N=%d: %d/%d converged from the reference start, %d from zero within 30 iterations" %
          (N, (g["status"] == 0).sum(), B, (g0["status"] == 0).sum()))
    s.close()
s = capi.Solver(8)
g = GPSRefTrajectory(mat_filename=1)
s.set_path(0, g.trajectory)
pose0 = np.tile(g.trajectory[0, [4, 5, 3]], (8, 1)) + 0.1
out = s.rollout(pose0, np.zeros(8, dtype=np.int32), 5)
print("rollout: status", out["log"][:, :, 6].ravel()[:8])
