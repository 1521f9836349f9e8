"""Golden vectors produced by RUNNING the reference's own Python code (tools/make_golden.py, build
container only) pin the two Python pieces either side of the solver: the reference generator
(ref_gps_traj.py:60-218) and the plant (vehicle_simulator.py:58-112).  Checked here against both the
host-side restatements in the package and the C restatements in the oracle.  Bit-exact where the
arithmetic is the same sequence of IEEE operations; 1e-12 where libm's trig differs from numpy's."""
import os

import numpy as np
import pytest

from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
from mkz_mpc_path_follower_b200.vehicle_simulator import VehicleSimulator

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def wp():
    return np.load(os.path.join(GOLD, "waypoints.npz"))


@pytest.mark.parametrize("pid", [1, 2, 3])
def test_trajectory_table(wp, pid):
    g = GPSRefTrajectory(mat_filename=pid)
    rows = wp["traj_p%d_rows" % pid]
    assert np.array_equal(g.trajectory[rows], wp["traj_p%d" % pid])          # bit-exact
    assert np.allclose(g.trajectory.sum(axis=0), wp["traj_p%d_colsum" % pid], rtol=1e-15, atol=0)
    # the stored x,y of the recording are the same projection
    from mkz_mpc_path_follower_b200 import paths
    d = paths.load_path(pid)
    assert np.abs(g.trajectory[:, 4] - d["x"]).max() < 1e-12 and np.abs(g.trajectory[:, 5] - d["y"]).max() < 1e-12


@pytest.mark.parametrize("pid", [1, 2, 3])
@pytest.mark.parametrize("H", [8, 20])
def test_get_waypoints_matches_reference(wp, oracle, pid, H):
    g = GPSRefTrajectory(mat_filename=pid, traj_horizon=H, traj_dt=0.2)
    key = "p%d_h%d" % (pid, H)
    q = wp[key + "_query"]
    path, keep = oracle.make_path(g.trajectory)
    n_fix = 0
    for r in range(q.shape[0]):
        X, Y, yaw = q[r]
        # time mode (track_using_time = True, the launch default)
        xr, yr, pr, stop = g.get_waypoints(X, Y, yaw)
        gold = wp[key + "_time_ref"][r]
        assert np.array_equal(np.stack((xr, yr, pr)), gold) and bool(stop) == bool(wp[key + "_time_stop"][r])
        ref_c, stop_c = oracle.get_waypoints(path, H, 0.2, X, Y, yaw)
        assert np.abs(ref_c - gold).max() <= 1e-9 and stop_c == bool(stop)
        n_fix += int(np.abs(gold[2] - yaw).max() >= np.pi or np.abs(np.diff(gold[2])).max() >= np.pi)
        # distance mode
        vt = float(wp[key + "_dist_v"][r])
        xr, yr, pr, stop = g.get_waypoints(X, Y, yaw, vt)
        gold = wp[key + "_dist_ref"][r]
        assert np.array_equal(np.stack((xr, yr, pr)), gold) and bool(stop) == bool(wp[key + "_dist_stop"][r])
        ref_c, stop_c = oracle.get_waypoints(path, H, 0.2, X, Y, yaw, v_target=vt)
        assert np.abs(ref_c - gold).max() <= 1e-9 and stop_c == bool(stop)
    # batched form = the single-query form
    refb, stopb = g.get_waypoints_batch(q[:, 0], q[:, 1], q[:, 2])
    assert np.array_equal(refb, wp[key + "_time_ref"]) and np.array_equal(stopb, wp[key + "_time_stop"].astype(bool))
    assert wp[key + "_time_stop"].any()          # the end-of-path latch is exercised


def test_plant_matches_reference(oracle):
    pl = np.load(os.path.join(GOLD, "plant.npz"))
    for case in range(3):
        X0, Y0, P0 = pl["case%d_init" % case]
        cmds, gold = pl["case%d_cmds" % case], pl["case%d_states" % case]
        sim = VehicleSimulator(X0=X0, Y0=Y0, Psi0=P0)
        st_c = np.array([X0, Y0, P0, 0, 0, 0, 0, 0], dtype=np.float64)
        for t in range(cmds.shape[0]):
            sim.mpc_cmd(cmds[t, 0], cmds[t, 1])
            sim.update_vehicle_model()
            assert np.abs(sim.full_state() - gold[t]).max() <= 1e-12 * max(1.0, np.abs(gold[t]).max()), (case, t)
            st_c = oracle.plant_step(st_c, cmds[t, 0], cmds[t, 1])
            assert np.abs(st_c - gold[t]).max() <= 1e-10 * max(1.0, np.abs(gold[t]).max()), (case, t)
    # the braking case reaches standstill: vx clamped at 0, lateral states frozen (:84-92)
    assert pl["case2_states"][-1, 3] == 0.0 and pl["case2_states"][-1, 4] == 0.0


def test_batched_plant_equals_scalar():
    rng = np.random.default_rng(1)
    B = 5
    sims = [VehicleSimulator(X0=i, Y0=2 * i, Psi0=0.1 * i) for i in range(B)]
    bat = VehicleSimulator(X0=np.arange(B), Y0=2.0 * np.arange(B), Psi0=0.1 * np.arange(B), batch=B)
    for t in range(60):
        a = rng.uniform(-1, 1, B); d = rng.uniform(-0.3, 0.3, B)
        bat.mpc_cmd(a, d); bat.update_vehicle_model()
        for i, s in enumerate(sims):
            s.mpc_cmd(a[i], d[i]); s.update_vehicle_model()
    assert np.array_equal(bat.full_state(), np.stack([s.full_state() for s in sims]))


def test_frenet_reference_fit_matches_reference():
    """frenet_ref.get_reference_frenet against golden vectors produced by RUNNING the reference's own
    get_reference_frenet (scripts/sim_path_utils/nav_msgs_path_frenet.py:76-86; tools/make_golden_frenet.py) on 24
    windows of the recorded paths: curvature polynomial, start heading and the resampled cubic path."""
    from mkz_mpc_path_follower_b200 import frenet_ref
    g = np.load(os.path.join(GOLD, "frenet_ref.npz"))
    n = int(g["n_cases"])
    assert n == 24
    for c in range(n):
        K, psi0, xi, yi = frenet_ref.get_reference_frenet({"x": g["c%d_x" % c], "y": g["c%d_y" % c], "s": g["c%d_s" % c]})
        Kg = g["c%d_K" % c]
        sq = np.linspace(0.0, float(g["c%d_s" % c][-1]), 9)
        # same least-squares problems, same LAPACK: the polynomials agree to rounding (compared as functions of s:
        # the individual coefficients of an ill-conditioned cubic fit may differ in their last digits)
        assert np.abs(np.polyval(K, sq) - np.polyval(Kg, sq)).max() <= 1e-12 * max(1.0, np.abs(np.polyval(Kg, sq)).max())
        assert np.allclose(K, Kg, rtol=1e-7, atol=1e-13)
        assert abs(psi0 - float(g["c%d_psi0" % c])) <= 1e-13
        assert np.abs(xi - g["c%d_xi" % c]).max() <= 1e-10 and np.abs(yi - g["c%d_yi" % c]).max() <= 1e-10
    # the batched form the workload uses solves the same two fits with fixed matrices instead of np.polyfit
    for c in (0, 7, 19):
        s, x, y = g["c%d_s" % c], g["c%d_x" % c], g["c%d_y" % c]
        s_fit = np.arange(0.0, s[-1], 0.5)
        Kb, pb = frenet_ref.fit_windows(np.interp(s_fit, s, x)[None], np.interp(s_fit, s, y)[None], float(s[-1]))
        sq = np.linspace(0.0, float(s[-1]), 9)
        Kg = g["c%d_K" % c]
        assert np.abs(np.polyval(Kb[0], sq) - np.polyval(Kg, sq)).max() <= 1e-9 * max(1.0, np.abs(np.polyval(Kg, sq)).max())
        assert abs(pb[0] - float(g["c%d_psi0" % c])) <= 1e-11
