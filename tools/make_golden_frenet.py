#!/usr/bin/env python
"""Generate tests/golden/frenet_ref.npz by RUNNING the reference's own curvature fit (build container only).

scripts/sim_path_utils/nav_msgs_path_frenet.py is read from /root/reference at run time (never copied into this
repo).  Only its fitting functions are executed: the source is cut, in memory, in front of its plotting / __main__
part (whose mixed tab / space indentation does not compile under Python 3), and rospy / nav_msgs / matplotlib /
rosbag / nav_msgs_common are replaced by stub modules.  `get_reference_frenet(path)` (:76-86) is then called on
windows of the recorded paths, expressed in the frame of a perturbed vehicle pose like the node's vehicle-frame
`target_path` message.

Output (committed): tests/golden/frenet_ref.npz
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import make_golden as MG  # noqa: E402

REF = MG.REF


def load_fit_functions():
    MG.stub_modules()
    nav = types.ModuleType("nav_msgs"); navmsg = types.ModuleType("nav_msgs.msg"); navmsg.Path = object
    sys.modules["nav_msgs"] = nav; sys.modules["nav_msgs.msg"] = navmsg
    sys.modules["nav_msgs_common"] = types.ModuleType("nav_msgs_common")
    path = os.path.join(REF, "scripts/sim_path_utils/nav_msgs_path_frenet.py")
    src = open(path).read()
    cut = src.index("# TEST FUNCTIONS FOR DEBUGGING/VERIFICATION.")
    mod = types.ModuleType("nav_msgs_path_frenet")
    exec(compile(src[:cut], path, "exec"), mod.__dict__)
    return mod


def main():
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    ref = load_fit_functions()
    rng = np.random.default_rng(20261019)
    out = {}
    case = 0
    for pid in (1, 2, 3):
        tr = GPSRefTrajectory(mat_filename=pid).trajectory
        s_all = tr[:, 6]
        for _ in range(8):
            i = int(rng.integers(0, np.searchsorted(s_all, s_all[-1] - 60.0)))
            length = float(rng.uniform(25.0, 55.0))
            j = int(np.searchsorted(s_all, s_all[i] + length))
            step = int(rng.integers(5, 40))                       # waypoint spacing of the message
            sel = np.arange(i, j, step)
            X0 = tr[i, 4] + rng.normal(scale=0.3); Y0 = tr[i, 5] + rng.normal(scale=0.3); yaw = tr[i, 3] + rng.normal(scale=0.05)
            dx = tr[sel, 4] - X0; dy = tr[sel, 5] - Y0
            x = np.cos(yaw) * dx + np.sin(yaw) * dy; y = -np.sin(yaw) * dx + np.cos(yaw) * dy
            s = s_all[sel] - s_all[sel[0]]
            K, psi0, xi, yi = ref.get_reference_frenet({"x": x, "y": y, "s": s})
            out["c%d_x" % case] = x; out["c%d_y" % case] = y; out["c%d_s" % case] = s
            out["c%d_K" % case] = np.asarray(K, dtype=np.float64); out["c%d_psi0" % case] = np.float64(psi0)
            out["c%d_xi" % case] = np.asarray(xi, dtype=np.float64); out["c%d_yi" % case] = np.asarray(yi, dtype=np.float64)
            case += 1
    out["n_cases"] = np.int64(case)
    f = os.path.join(MG.OUT, "frenet_ref.npz")
    np.savez_compressed(f, **out)
    print("frenet_ref.npz", os.path.getsize(f), "cases", case)


if __name__ == "__main__":
    main()
