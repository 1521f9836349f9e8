#!/usr/bin/env python
"""Single-solve latency on the reference workload (BASELINE.json configs[0]): the closed loop of
scripts/mpc_cmd_pub.jl:86-157 around the plant of scripts/vehicle_simulator.py, N = 8, dt = 0.2, one vehicle, from the
launch file's start pose on path 3 (launch/sim_path_follow.launch:13,23-25) and from the start of path 1, until the
stop latch.  Every control step is ONE call of mpcb200_solve_batch with host buffers and a batch of one (H2D + kernel +
D2H inside the timed call), warm-started from the previous solution like the reference.  The CPU oracle solves the same
sequence of problems (same states, references, previous commands, start points) beside it.
python tools/closed_loop_latency.py  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi  # noqa: E402
from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory  # noqa: E402
from mkz_mpc_path_follower_b200.vehicle_simulator import VehicleSimulator  # noqa: E402
from oracle import oracle as O  # noqa: E402  (the CPU baseline leg)


def drive(solver, ocfg, path_id, pose0, max_steps=800):
    N = solver.N
    g = GPSRefTrajectory(mat_filename=path_id, traj_horizon=N, traj_dt=0.2)
    sim = VehicleSimulator(X0=pose0[0], Y0=pose0[1], Psi0=pose0[2], batch=1)
    u_curr = np.zeros((1, 2)); warm = np.zeros((1, 6 * N + 4)); warm_o = np.zeros(6 * N + 4)
    gpu_ms, cpu_ms, iters, status, mism = [], [], [], [], 0
    for t in range(max_steps):
        for _ in range(10):
            sim.update_vehicle_model()
        st = sim.state_est()[:, :4].copy()
        ref, stop = g.get_waypoints_batch(st[:, 0], st[:, 1], st[:, 2])
        if stop[0]:
            break
        vdes = np.array([1.0])
        warm_o[:] = warm[0]
        t0 = time.perf_counter()
        out = solver.solve_batch(st, ref, u_curr, v_des=vdes, warm=warm)
        gpu_ms.append(1e3 * (time.perf_counter() - t0))
        t0 = time.perf_counter()
        ro = O.solve(ocfg, st[0], ref[0], 1.0, u_curr[0], warm=warm_o)
        cpu_ms.append(1e3 * (time.perf_counter() - t0))
        mism += int(ro["status"] != out["status"][0] or np.abs(ro["u0"] - out["u0"][0]).max() > 1e-5)
        iters.append(int(out["iters"][0])); status.append(int(out["status"][0]))
        sim.mpc_cmd(out["u0"][:, 0], out["u0"][:, 1])
        u_curr[0, 0] = out["u0"][0, 1]; u_curr[0, 1] = out["u0"][0, 0]
    return gpu_ms, cpu_ms, iters, status, mism


def main():
    N = 8
    solver = capi.Solver(N)
    O.build()
    ocfg = O.default_cfg(N, max_iter=int(solver.cfg.max_iter))
    gpu, cpu, its, stat, mism = [], [], [], [], 0
    g1 = GPSRefTrajectory(mat_filename=1).trajectory
    for path_id, pose0 in ((3, (0.0, 3.0, -1.5)), (1, (g1[0, 4], g1[0, 5], g1[0, 3]))):
        a, b, c, d, m = drive(solver, ocfg, path_id, pose0)
        gpu += a[5:]; cpu += b[5:]; its += c[5:]; stat += d[5:]; mism += m     # the first solves include one-time set-up
    gpu, cpu = np.array(gpu), np.array(cpu)
    print(json.dumps({"workload": "configs[0]: closed loop, N=8, path3 from the launch pose + path1 from its start, until stop_cmd",
                      "solves": int(gpu.size), "optimal_frac": float(np.mean(np.array(stat) == 0)), "mean_iters": float(np.mean(its)),
                      "gpu_c_abi_ms": {"p50": float(np.percentile(gpu, 50)), "p90": float(np.percentile(gpu, 90)), "p99": float(np.percentile(gpu, 99))},
                      "cpu_oracle_ms": {"p50": float(np.percentile(cpu, 50)), "p90": float(np.percentile(cpu, 90)), "p99": float(np.percentile(cpu, 99)),
                                        "kind": "port (restated interior point, NOT Ipopt), 1 thread"},
                      "mismatches_vs_oracle": int(mism)}))


if __name__ == "__main__":
    main()
