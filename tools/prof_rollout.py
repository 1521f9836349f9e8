#!/usr/bin/env python
"""Small closed-loop rollout for ncu: python tools/prof_rollout.py [vehicles] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mkz_mpc_path_follower_b200 import capi  # noqa: E402
from rollout_bench import fleet  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
trajs, path_of, pose0 = fleet(V)
s = capi.Solver(8)
for i, g in enumerate(trajs):
    s.set_path(i, g.trajectory)
for r in range(2):
    out = s.rollout(pose0, path_of, T)
    print("rep %d: kernel %.3f ms, optimal %.4f, mean iters %.2f" % (r, s.stats()["kernel_ms"], (out["log"][:, :, 6] == 0).mean(), out["log"][:, :, 7].mean()))
