#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING the reference's own Python code (build container only).

The reference's ref_gps_traj.py and vehicle_simulator.py are Python 2 ROS nodes.  They are read
from /root/reference at run time (never copied into this repo), their two Python-2 `print`
statements are rewritten in memory, rospy / rosbag / matplotlib / the ROS message package are
replaced by stub modules, and the classes are then exercised exactly as the nodes would:

  * GPSRefTrajectory(...).get_waypoints(X, Y, yaw[, v_target])   (ref_gps_traj.py:131-218)
  * VehicleSimulator._update_vehicle_model()                      (vehicle_simulator.py:58-112)

Outputs (committed):  tests/golden/waypoints.npz, tests/golden/plant.npz
"""
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PARAMS = {"lat0": 37.917929, "lon0": -122.331798, "yaw0": 0.0, "is_heading_info": True,
          "X0": 0.0, "Y0": 3.0, "Psi0": -1.5}


def stub_modules():
    rospy = types.ModuleType("rospy")
    rospy.has_param = lambda k: k in PARAMS
    rospy.get_param = lambda k, d=None: PARAMS.get(k, d)
    rospy.init_node = lambda *a, **k: None
    rospy.Subscriber = lambda *a, **k: None
    rospy.Publisher = lambda *a, **k: types.SimpleNamespace(publish=lambda m: None)
    rospy.Rate = lambda hz: types.SimpleNamespace(sleep=lambda: None)
    rospy.is_shutdown = lambda: True
    rospy.Time = types.SimpleNamespace(now=lambda: 0.0)
    rospy.ROSInterruptException = Exception
    sys.modules["rospy"] = rospy
    sys.modules["rosbag"] = types.ModuleType("rosbag")
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    pkg = types.ModuleType("mkz_mpc_path_follower"); msg = types.ModuleType("mkz_mpc_path_follower.msg")
    msg.state_est = lambda: types.SimpleNamespace(header=types.SimpleNamespace(stamp=None))
    msg.MPC_cmd = lambda: types.SimpleNamespace()
    sys.modules["mkz_mpc_path_follower"] = pkg; sys.modules["mkz_mpc_path_follower.msg"] = msg


def load_py2(path, name):
    src = open(path).read()
    src = re.sub(r"^(\s*)print\s+'([^']*)'\s*$", r"\1print('\2')", src, flags=re.M)
    # the other Python-2-ism: `1/m` with the int m = 1840 (vehicle_simulator.py:86) is INTEGER division there
    # (= 0: the tyre-force term drops out of the vx equation as the node runs); keep that under Python 3
    n_div = src.count("1/m*")
    src = src.replace("1/m*", "(1//m)*")
    if name == "vehicle_simulator":
        assert n_div == 1, "expected exactly one int/int division in vehicle_simulator.py"

    mod = types.ModuleType(name)
    mod.__dict__["__name__"] = name
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def main():
    os.makedirs(OUT, exist_ok=True)
    stub_modules()
    rgt = load_py2(os.path.join(REF, "scripts/gps_utils/ref_gps_traj.py"), "ref_gps_traj")
    rng = np.random.default_rng(20261018)
    out = {}
    for pid in (1, 2, 3):
        for (H, dt) in ((8, 0.2), (20, 0.2)):
            g = rgt.GPSRefTrajectory(mat_filename=os.path.join(REF, "paths/path%d_6_20.mat" % pid), traj_horizon=H, traj_dt=dt)
            tr = g.get_global_trajectory_reference()
            n = tr.shape[0]
            if H == 8:   # the (n,7) table of ref_gps_traj.py:106: every 16th row, the last row and column sums
                out["traj_p%d_rows" % pid] = np.concatenate((np.arange(0, n, 16), [n - 1]))
                out["traj_p%d" % pid] = tr[out["traj_p%d_rows" % pid]].copy()
                out["traj_p%d_colsum" % pid] = tr.sum(axis=0)
            # queries: along the path (incl. the very start / very end -> stop_cmd, and the psi wrap samples)
            jumps = np.nonzero(np.abs(np.diff(tr[:, 3])) > np.pi)[0]
            idx = np.concatenate((rng.integers(0, n, size=40), [0, 1, n - 1, n - 30, n - 200], jumps, np.maximum(jumps - 40, 0)))
            q = np.empty((len(idx), 3)); res_t = []; res_d = []; stop_t = []; stop_d = []; vts = []
            for r, i in enumerate(idx):
                X = tr[i, 4] + rng.normal(scale=0.5); Y = tr[i, 5] + rng.normal(scale=0.5)
                yaw = tr[i, 3] + rng.normal(scale=0.1)
                if r % 7 == 3:
                    yaw += 2 * np.pi * rng.choice([-1, 1])   # forces the wrap-around fix
                q[r] = (X, Y, yaw)
                xr, yr, pr, st = g.get_waypoints(X, Y, yaw)                 # time mode
                res_t.append(np.stack((xr, yr, np.array(pr, copy=True)))); stop_t.append(st)
                vt = 1.0 + 9.0 * rng.random()
                vts.append(vt)
                xr, yr, pr, st = g.get_waypoints(X, Y, yaw, vt)              # distance mode
                res_d.append(np.stack((xr, yr, np.array(pr, copy=True)))); stop_d.append(st)
            key = "p%d_h%d" % (pid, H)
            out[key + "_query"] = q
            out[key + "_time_ref"] = np.array(res_t); out[key + "_time_stop"] = np.array(stop_t)
            out[key + "_dist_ref"] = np.array(res_d); out[key + "_dist_stop"] = np.array(stop_d)
            out[key + "_dist_v"] = np.array(vts)
    np.savez_compressed(os.path.join(OUT, "waypoints.npz"), **out)

    # ---- plant: drive VehicleSimulator._update_vehicle_model with recorded command sequences
    vs = load_py2(os.path.join(REF, "scripts/vehicle_simulator.py"), "vehicle_simulator")
    cls = vs.VehicleSimulator
    cls.pub_loop = lambda self: None          # __init__ calls pub_loop(); the test steps the model itself
    plant = {}
    for case, (X0, Y0, P0) in enumerate(((0.0, 3.0, -1.5), (142.0, -82.0, 2.0), (-300.0, -450.0, 1.0))):
        PARAMS.update({"X0": X0, "Y0": Y0, "Psi0": P0})
        sim = cls()
        T = 600
        cmds = np.zeros((T, 2)); states = np.zeros((T, 8))
        a, d = 0.0, 0.0
        for t in range(T):
            if t % 10 == 0:   # a new command every control period
                a = float(np.clip(a + rng.normal(scale=0.4), -1.0, 1.0)) if t > 0 else 1.0
                d = float(np.clip(d + rng.normal(scale=0.05), -0.5, 0.5))
                if case == 2 and t > 300:
                    a = -1.0      # brake to standstill: exercises the vx clamp and the frozen lateral states
            sim.acc_des, sim.df_des = a, d
            sim._update_vehicle_model()
            cmds[t] = (a, d)
            states[t] = (sim.X, sim.Y, sim.psi, sim.vx, sim.vy, sim.wz, sim.acc, sim.df)
        plant["case%d_init" % case] = np.array([X0, Y0, P0])
        plant["case%d_cmds" % case] = cmds
        plant["case%d_states" % case] = states
    np.savez_compressed(os.path.join(OUT, "plant.npz"), **plant)
    for f in ("waypoints.npz", "plant.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
