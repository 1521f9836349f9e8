"""CPU check of the CUDA solver's source: mpc_kernel.cuh compiled by g++ against a 32-lane
coroutine emulator (tests/emu, test infrastructure only) must follow the oracle iterate for
iterate.  This is what lets kernel changes be checked without a GPU; the real parity tests are
the -m gpu ones, through the C ABI."""
import os
import sys

import numpy as np
import pytest

from mkz_mpc_path_follower_b200 import workload as W

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


@pytest.fixture(autouse=True)
def _race_check():
    """Every emulated run is also a shared-memory race check (compute-sanitizer's racecheck is closed on
    the GPU pool): a word written by one lane may be read or overwritten by another only across a
    bar.warp.sync / block barrier."""
    import emu as E
    E.race_check(True)
    yield
    assert E.race_count() == 0


@pytest.mark.parametrize("N,B", [(8, 24), (20, 10), (3, 4)])
def test_emulated_kernel_matches_oracle(oracle, N, B):
    import emu as E
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=4)
    e = E.solve_batch(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True)
    assert (o["status"] == e["status"]).all()
    ok = o["status"] == 0
    assert ok.sum() >= B // 2
    assert np.abs(o["u0"] - e["u0"])[ok].max() <= 1e-9
    assert (np.abs(o["cost"] - e["cost"])[ok] <= 1e-9 * np.maximum(1, np.abs(o["cost"][ok]))).all()
    assert np.abs(o["traj"] - e["traj"])[ok].max() <= 1e-8
    # warm start path
    warm_o = o["traj"].copy(); warm_e = o["traj"].copy()
    o2 = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=warm_o, n_threads=4)
    e2 = E.solve_batch(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"], warm=warm_e)
    assert (o2["status"] == e2["status"]).all() and (o2["iters"] == e2["iters"]).all()
    assert np.abs(o2["u0"] - e2["u0"])[o2["status"] == 0].max() <= 1e-9


@pytest.mark.parametrize("N,B", [(32, 3), (40, 6), (63, 3), (64, 3), (80, 4)])
def test_emulated_long_horizon_team(oracle, N, B):
    """N > 31: one block of 2 (N <= 63) or 3 warps per problem; the emulator steps all 64/96 threads,
    bar.sync included.  Start = the reference waypoints (the all-zero start is hundreds of metres
    away and rarely converges within the cap at these horizons)."""
    import emu as E
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    w0 = W.reference_start(b, N)
    wo, we = w0.copy(), w0.copy()
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=wo, n_threads=4)
    e = E.solve_batch(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"], warm=we)
    assert (o["status"] == 0).all() and (e["status"] == 0).all()
    assert (o["iters"] == e["iters"]).all()
    assert np.abs(o["u0"] - e["u0"]).max() <= 1e-9
    assert (np.abs(o["cost"] - e["cost"]) <= 1e-9 * np.maximum(1, np.abs(o["cost"]))).all()
    assert np.abs(wo - we).max() <= 1e-8   # the whole returned trajectory


def test_emulated_long_horizon_zero_start(oracle):
    import emu as E
    N, B = 40, 3
    cfg = oracle.default_cfg(N, max_iter=120)
    b = W.make_batch(B, N)
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=4)
    e = E.solve_batch(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"])
    assert (o["status"] == e["status"]).all() and (o["iters"] == e["iters"]).all()
    ok = o["status"] == 0
    assert ok.any()
    assert np.abs(o["u0"] - e["u0"])[ok].max() <= 1e-9


def test_restoration_by_rollout(oracle, monkeypatch):
    """Problems whose filter line search fails from the all-zero start (Ipopt would enter its restoration
    phase): oracle and emulated kernel must restore identically (same iteration counts), converge, and
    end at the optimum that the reference-waypoint start finds."""
    import emu as E
    N = 20
    idx = np.array([38, 62, 79, 122, 163])
    b = W.make_batch(int(idx.max()) + 1, N)
    sel = {k: b[k][idx] for k in ("state", "ref", "v_des", "u_prev")}
    cfg = oracle.default_cfg(N)
    monkeypatch.setenv("MPC_ORACLE_NO_RESTO", "1")
    o0 = oracle.solve_batch(cfg, sel["state"], sel["ref"], sel["v_des"], sel["u_prev"], n_threads=4)
    assert (o0["status"] == 4).all()            # without restoration: Error
    monkeypatch.delenv("MPC_ORACLE_NO_RESTO")
    o = oracle.solve_batch(cfg, sel["state"], sel["ref"], sel["v_des"], sel["u_prev"], n_threads=4)
    e = E.solve_batch(E.kcfg_from_oracle(cfg), sel["state"], sel["ref"], sel["v_des"], sel["u_prev"])
    assert (o["status"] == 0).all() and (e["status"] == 0).all()
    assert (o["iters"] == e["iters"]).all()
    assert np.abs(o["u0"] - e["u0"]).max() <= 1e-9
    w = W.reference_start(sel, N)
    o2 = oracle.solve_batch(cfg, sel["state"], sel["ref"], sel["v_des"], sel["u_prev"], warm=w, n_threads=4)
    assert (o2["status"] == 0).all()
    assert np.abs(o["u0"] - o2["u0"]).max() <= 1e-5
    assert (np.abs(o["cost"] - o2["cost"]) <= 1e-6 * np.maximum(1, o2["cost"])).all()


@pytest.mark.parametrize("N,B", [(8, 16), (20, 12), (40, 3)])
def test_rollout_start_mode(oracle, N, B):
    """MPCB200_START_ROLLOUT (opt-in): previous command held over the horizon, model rolled out from the
    measured state.  The emulated kernel must match the oracle started from mpc_oracle_rollout_start,
    converge in a handful of iterations, and reach the optimum of the all-zero start."""
    import emu as E
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    w = oracle.rollout_start(cfg, b["state"], b["u_prev"])
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=w.copy(), n_threads=4)
    e = E.solve_batch(E.kcfg_from_oracle(cfg, start_mode=1), b["state"], b["ref"], b["v_des"], b["u_prev"])
    assert (o["status"] == 0).all() and (e["status"] == 0).all()
    assert (o["iters"] == e["iters"]).all() and o["iters"].mean() < 15
    assert np.abs(o["u0"] - e["u0"]).max() <= 1e-9
    o0 = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=4)
    both = o0["status"] == 0
    assert np.abs(o["u0"] - o0["u0"])[both].max() <= 1e-5


def test_emulated_rollout_group(oracle):
    """The closed-loop rollout kernel source on the emulator: one block of four warps = four vehicles, the
    plant of all of them integrated by one warp (lane = vehicle) between two block barriers per control
    period.  Five vehicles (a full group and a ragged one) against the oracle's closed loop."""
    import emu as E
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    g = GPSRefTrajectory(mat_filename=1)
    cfg = oracle.default_cfg(8)
    rng = np.random.default_rng(3)
    poses = np.array([g.trajectory[j, [4, 5, 3]] + rng.normal(scale=[0.3, 0.3, 0.03]) for j in (0, 900, 2500, 4000, 5200)])
    T = 10
    seed = oracle.module_load_solution(cfg)    # the first solve starts from the module-load solution (MKZMPCPathFollower.jl:126-128)
    log, final = E.rollout(E.kcfg_from_oracle(cfg), g.trajectory, poses, T, warm0=seed)
    path, keep = oracle.make_path(g.trajectory)
    for b in range(poses.shape[0]):
        olog = oracle.closed_loop(cfg, path, poses[b], T)
        assert np.array_equal(log[:, b, 6], olog[:, 6]) and np.array_equal(log[:, b, 7], olog[:, 7]), b
        assert np.abs(log[:, b, 4:6] - olog[:, 4:6]).max() <= 1e-9, b
        assert np.abs(log[:, b, 0:4] - olog[:, 0:4]).max() <= 1e-9, b


def test_emulated_long_horizon_rollout_and_on_path(oracle):
    """N = 40: one block of two warps per vehicle.  The team-wide get_waypoints (nearest-sample arg-min reduced across
    the warps, neighbour exchange for the heading unwrap) against the host generator bit for bit, and three closed-loop
    steps of the long-horizon rollout against the oracle's closed loop."""
    import emu as E
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N = 40
    g = GPSRefTrajectory(mat_filename=2, traj_horizon=N, traj_dt=0.2)
    cfg = oracle.default_cfg(N, max_iter=60)
    rng = np.random.default_rng(5)
    n = g.trajectory.shape[0]
    jumps = np.nonzero(np.abs(np.diff(g.trajectory[:, 3])) > np.pi)[0]
    idx = [100, 2000, int(jumps[0]) - 300 if len(jumps) else 3000, n - 50]
    state = np.array([[g.trajectory[j, 4] + rng.normal(scale=0.3), g.trajectory[j, 5] + rng.normal(scale=0.3),
                       g.trajectory[j, 3] + rng.normal(scale=0.03), 5.0] for j in idx])
    state[1, 2] += 2 * np.pi    # forces the wrap-around fix
    for mode_time, vt in ((True, 1.0), (False, 6.5)):
        e = E.solve_batch_on_path(E.kcfg_from_oracle(cfg), g.trajectory, state, np.zeros((len(idx), 2)), track_using_time=mode_time, target_vel=vt)
        ref, stop = g.get_waypoints_batch(state[:, 0], state[:, 1], state[:, 2], v_target=None if mode_time else vt)
        assert np.array_equal(e["ref"], ref)          # bit for bit: every operation rounded like numpy's
        assert np.array_equal(e["stop"].astype(bool), stop)
    assert E.race_count() == 0
    # closed loop, two vehicles, three control steps
    cfg = oracle.default_cfg(N, max_iter=200)
    poses = np.array([g.trajectory[j, [4, 5, 3]] + rng.normal(scale=[0.2, 0.2, 0.02]) for j in (500, 3000)])
    seed = oracle.module_load_solution(cfg)
    log, final = E.rollout(E.kcfg_from_oracle(cfg), g.trajectory, poses, 3, warm0=seed)
    path, keep = oracle.make_path(g.trajectory)
    for b in range(2):
        olog = oracle.closed_loop(cfg, path, poses[b], 3)
        assert np.array_equal(log[:, b, 6], olog[:, 6]), (b, log[:, b, 6:8], olog[:, 6:8])
        assert np.abs(log[:, b, 7] - olog[:, 7]).max() <= 2
        assert np.abs(log[:, b, 4:6] - olog[:, 4:6]).max() <= 1e-6, b
        assert np.abs(log[:, b, 0:4] - olog[:, 0:4]).max() <= 1e-9, b


def test_emulated_on_path_bad_path_id(oracle):
    """A path id outside the tables (possible only with device pointers, which the host cannot inspect) is answered with
    status Error and zero commands instead of an out-of-range table access."""
    import emu as E
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    g = GPSRefTrajectory(mat_filename=1)
    cfg = oracle.default_cfg(8)
    state = np.array([[g.trajectory[50, 4], g.trajectory[50, 5], g.trajectory[50, 3], 3.0]] * 3)
    e = E.solve_batch_on_path(E.kcfg_from_oracle(cfg), g.trajectory, state, np.zeros((3, 2)), path_of=np.array([0, 7, -1]))
    assert e["status"].tolist() == [0, 4, 4] and np.all(e["u0"][1:] == 0.0) and e["iters"][1:].tolist() == [0, 0]


def test_line_search_failure_at_an_acceptable_point(oracle):
    """A problem of the 4-GPU sweep (rollout start, N = 20) whose line search fails on rounding alone two
    iterations before the tolerance would be met (theta ~ 4e-13, step 5e-7).  Ipopt returns such a point
    ("restoration phase called at almost feasible / acceptable point" -> Solved_To_Acceptable_Level);
    before that rule the restoration by rollout threw the iterate away and the solve ran to the cap
    (one straggler doubled a rank's kernel time).  The two linear solvers round differently here, so
    the iteration counts may differ by a couple; the returned point may not."""
    import emu as E
    N = 20
    b = W.make_batch(1, N, b0=131072 + 127164)
    cfg = oracle.default_cfg(N)
    w = oracle.rollout_start(cfg, b["state"], b["u_prev"])
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=w.copy(), n_threads=1)
    e = E.solve_batch(E.kcfg_from_oracle(cfg, start_mode=1), b["state"], b["ref"], b["v_des"], b["u_prev"])
    assert o["status"][0] == 0 and e["status"][0] == 0
    assert o["iters"][0] <= 20 and e["iters"][0] <= 20
    assert np.abs(o["u0"] - e["u0"]).max() <= 1e-7
    o0 = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=1)   # all-zero start
    assert o0["status"][0] == 0 and np.abs(o["u0"] - o0["u0"]).max() <= 1e-6
    assert abs(o["cost"][0] - o0["cost"][0]) <= 1e-6 * o0["cost"][0]


def test_emulated_kernel_property(oracle):
    """Property-based: problems of the synthetic stream with random extra offsets, speeds and previous commands,
    started from the reference waypoints -- emulated CUDA source and oracle end with the same status and, where
    that is Optimal, the same command."""
    import emu as E
    from hypothesis import given, settings, strategies as st
    N = 5
    cfg = oracle.default_cfg(N, max_iter=80)
    kc = E.kcfg_from_oracle(cfg)
    fl = lambda lo, hi: st.floats(min_value=lo, max_value=hi, allow_nan=False, allow_infinity=False)

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(st.integers(0, 10 ** 6), fl(-2.0, 2.0), fl(-2.0, 2.0), fl(-0.3, 0.3), fl(0.0, 20.0), fl(-0.5, 0.5), fl(-1.0, 1.0))
    def check(b0, dx, dy, dpsi, v, df0, a0):
        b = W.make_batch(1, N, b0=b0)
        b["state"][0] += [dx, dy, dpsi, 0.0]; b["state"][0, 3] = v
        b["u_prev"][0] = [df0, a0]
        w = W.reference_start(b, N)
        o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=w.copy(), n_threads=1)
        e = E.solve_batch(kc, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=w.copy())
        assert o["status"][0] == e["status"][0]
        if o["status"][0] == 0:
            assert abs(int(o["iters"][0]) - int(e["iters"][0])) <= 1
            assert np.abs(o["u0"] - e["u0"]).max() <= 1e-7

    check()


@pytest.mark.parametrize("N", [3, 40])
def test_emulated_nearest_sample_search(oracle, N):
    """TeamSolver::nearest_sample skips the chunks of path samples whose bounding circle cannot hold the nearest one; the result
    must still be np.argmin over the whole path (ref_gps_traj.py:136-137), so the generated waypoints stay bit-equal to the host
    generator's: poses on the samples, between two samples (near-ties across chunk borders), metres to tens of kilometres
    away, at the centroid of the path, and not finite (argmin of an all-NaN array is 0, not an out-of-range index)."""
    import emu as E
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    k = E.kcfg_from_oracle(oracle.default_cfg(N, max_iter=1))
    rng = np.random.default_rng(3)
    m = 12 if N > 31 else 40
    for pid in (1, 2, 3):
        g = GPSRefTrajectory(mat_filename=pid, traj_horizon=N, traj_dt=0.2)
        tr = g.trajectory
        n = tr.shape[0]
        poses = []
        for scale in (0.0, 1e-9, 0.3, 5.0, 50.0, 500.0, 5e4):
            j = rng.integers(0, n, m)
            poses.append(np.stack([tr[j, 4] + rng.normal(0, 1, m) * scale, tr[j, 5] + rng.normal(0, 1, m) * scale,
                                   tr[j, 3] + rng.normal(0, 0.1, m), np.full(m, 5.0)], 1))
        j = np.arange(14, n - 2, 16 * 13)[:m]      # sample 15 | 16 is a chunk border
        for o in (0, 1):
            poses.append(np.stack([0.5 * (tr[j + o, 4] + tr[j + o + 1, 4]), 0.5 * (tr[j + o, 5] + tr[j + o + 1, 5]), tr[j, 3], np.full(len(j), 5.0)], 1))
        poses.append(np.array([[tr[:, 4].mean(), tr[:, 5].mean(), 0.0, 5.0], [tr[:, 4].min() - 10, tr[:, 5].max() + 10, 1.0, 5.0],
                               [np.nan, 0.0, 0.0, 5.0], [0.0, np.inf, 0.0, 5.0], [1e200, 1e200, 0.0, 5.0]]))
        st = np.concatenate(poses)
        for mode_time, vt in ((True, 1.0), (False, 6.5)):
            e = E.solve_batch_on_path(k, tr, st, np.zeros((len(st), 2)), track_using_time=mode_time, target_vel=vt)
            with np.errstate(all="ignore"):
                ref, stop = g.get_waypoints_batch(st[:, 0], st[:, 1], st[:, 2], v_target=None if mode_time else vt)
            assert np.array_equal(e["ref"], ref, equal_nan=True)
            assert np.array_equal(e["stop"].astype(bool), stop)
