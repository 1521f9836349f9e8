"""CPU check of the thread-per-problem solver's source (csrc/tpp_solver.cuh, the large-batch layout): it is plain
scalar code, so tests/emu compiles it with g++ as it stands and runs it against the oracle.  It has to follow the oracle
iterate for iterate like the warp-per-problem kernel does (same iteration, only the linear algebra differs from the
oracle's).  The -m gpu tests drive the same source through the C ABI."""
import os
import sys

import numpy as np
import pytest

from mkz_mpc_path_follower_b200 import workload as W

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N,B", [(8, 256), (20, 192), (3, 64), (31, 24)])
def test_thread_per_problem_matches_oracle(oracle, N, B):
    import emu as E
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=4)
    e = E.solve_batch_tpp(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, slots=37)
    assert (o["status"] == e["status"]).all()
    # (the different rounding of Riccati vs dense LDL' shows after ~100 iterations, the normal case from the zero start at N = 31)
    assert (o["iters"] == e["iters"]).mean() >= (0.99 if N <= 20 else 0.85)
    assert (o["n_resto"] == e["n_resto"]).all()
    ok = o["status"] == 0
    assert ok.sum() >= B // 2
    assert np.abs(o["u0"] - e["u0"])[ok].max() <= 1e-8
    assert (np.abs(o["cost"] - e["cost"])[ok] <= 1e-8 * np.maximum(1, np.abs(o["cost"][ok]))).all()
    assert np.abs(o["traj"] - e["traj"])[ok].max() <= 1e-6
    # warm start: in/out buffer, a handful of iterations
    wo, we = o["traj"].copy(), o["traj"].copy()
    o2 = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=wo, n_threads=4)
    e2 = E.solve_batch_tpp(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"], warm=we, slots=5)
    assert (o2["status"] == e2["status"]).all() and (o2["iters"] == e2["iters"]).all()
    assert np.abs(o2["u0"] - e2["u0"])[o2["status"] == 0].max() <= 1e-9
    assert np.abs(wo - we)[o2["status"] == 0].max() <= 1e-8


def test_thread_per_problem_equals_the_warp_kernel_source(oracle):
    """The two device layouts of the solver, both emulated: same statuses and iteration counts, results equal to rounding."""
    import emu as E
    N, B = 8, 48
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    k = E.kcfg_from_oracle(cfg)
    w = E.solve_batch(k, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True)
    t = E.solve_batch_tpp(k, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True)
    assert (w["status"] == t["status"]).all() and (w["iters"] == t["iters"]).all()
    assert np.abs(w["traj"] - t["traj"]).max() <= 1e-9 and np.abs(w["cost"] - t["cost"]).max() <= 1e-9


@pytest.mark.parametrize("N", [40, 80])
def test_thread_per_problem_long_horizons(oracle, N):
    """Any horizon: the passes are loops over the stages (no team size, no shared-memory budget)."""
    import emu as E
    B = 6
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    w0 = W.reference_start(b, N)
    wo, we = w0.copy(), w0.copy()
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=wo, n_threads=4)
    e = E.solve_batch_tpp(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"], warm=we)
    assert (o["status"] == 0).all() and (e["status"] == 0).all() and (o["iters"] == e["iters"]).all()
    assert np.abs(o["u0"] - e["u0"]).max() <= 1e-9 and np.abs(wo - we).max() <= 1e-7


def test_thread_per_problem_edge_cases(oracle):
    """Infeasible initial speed / previous command (decided a priori), iteration cap, rollout start, a restored problem."""
    import emu as E
    N = 8
    cfg = oracle.default_cfg(N)
    b = W.make_batch(8, N)
    st = b["state"].copy(); up = b["u_prev"].copy()
    st[0, 3] = 25.0          # v0 > v_max
    st[1, 3] = -1.0          # v0 < v_min
    up[2, 0] = 0.9           # previous steering further from the box than one rate step
    o = oracle.solve_batch(cfg, st, b["ref"], b["v_des"], up, n_threads=2)
    e = E.solve_batch_tpp(E.kcfg_from_oracle(cfg), st, b["ref"], b["v_des"], up)
    assert (o["status"] == e["status"]).all() and (e["status"][:3] == 1).all() and (e["iters"][:3] == 0).all()
    cfg2 = oracle.default_cfg(N, max_iter=7)
    o = oracle.solve_batch(cfg2, b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=2)
    e = E.solve_batch_tpp(E.kcfg_from_oracle(cfg2), b["state"], b["ref"], b["v_des"], b["u_prev"])
    assert (o["status"] == e["status"]).all() and (e["status"] == 3).all() and (e["iters"] == 7).all()
    assert np.abs(o["u0"] - e["u0"]).max() <= 1e-9
    # MPCB200_START_ROLLOUT
    k = E.kcfg_from_oracle(cfg, start_mode=1)
    w = E.solve_batch(k, b["state"], b["ref"], b["v_des"], b["u_prev"])
    t = E.solve_batch_tpp(k, b["state"], b["ref"], b["v_des"], b["u_prev"])
    assert (w["status"] == t["status"]).all() and (w["iters"] == t["iters"]).all() and np.abs(w["u0"] - t["u0"]).max() <= 1e-9
    # restorations by rollout happen in this batch and are counted alike
    N = 20
    cfg = oracle.default_cfg(N)
    b = W.make_batch(256, N)
    o = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=4)
    e = E.solve_batch_tpp(E.kcfg_from_oracle(cfg), b["state"], b["ref"], b["v_des"], b["u_prev"])
    assert o["n_resto"].sum() > 0 and (o["n_resto"] == e["n_resto"]).all()
    r = o["n_resto"] > 0
    assert (o["status"][r] == e["status"][r]).all() and np.abs(o["u0"] - e["u0"])[r & (o["status"] == 0)].max() <= 1e-7


def test_thread_per_problem_closed_loop(oracle):
    """The closed loop as mpcb200_rollout runs it for large fleets: per control period the plant, get_waypoints written into the
    vehicle's slot of the solver state, and the thread-per-problem solve warm-started from what the slot still holds.  Against
    the oracle's closed loop (plant, reference generation, warm-started solves, command feedback, stop latch), a vehicle that
    reaches the end of the path included, and against the fused rollout kernel's source."""
    import emu as E
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    g = GPSRefTrajectory(mat_filename=1)
    cfg = oracle.default_cfg(8)
    rng = np.random.default_rng(3)
    n = g.trajectory.shape[0]
    poses = np.array([g.trajectory[j, [4, 5, 3]] + rng.normal(scale=[0.3, 0.3, 0.03]) for j in (0, 900, 2500, 4000, 5200, n - 40)])
    T = 12
    seed = oracle.module_load_solution(cfg)
    log, final = E.rollout(E.kcfg_from_oracle(cfg), g.trajectory, poses, T, warm0=seed, tpp=True)
    path, keep = oracle.make_path(g.trajectory)
    for b in range(poses.shape[0]):
        olog = oracle.closed_loop(cfg, path, poses[b], T)
        assert np.array_equal(log[:, b, 6], olog[:, 6]) and np.array_equal(log[:, b, 7], olog[:, 7]), b
        assert np.abs(log[:, b, 4:6] - olog[:, 4:6]).max() <= 1e-9, b
        assert np.abs(log[:, b, 0:4] - olog[:, 0:4]).max() <= 1e-9, b
    assert (log[:, -1, 6] == -1).any() and (log[-1, -1, 4:6] == [-1.0, 0.0]).all()     # the last vehicle ran into the stop latch
    wlog, wfinal = E.rollout(E.kcfg_from_oracle(cfg), g.trajectory, poses, T, warm0=seed)
    assert np.array_equal(log[:, :, 6:8], wlog[:, :, 6:8]) and np.abs(log - wlog).max() <= 1e-9 and np.abs(final - wfinal).max() <= 1e-9


def _frenet_stress_batch(B, N, seed):
    """Tighter curves and larger offsets than the recorded paths give (as in tests/test_frenet.py)."""
    rng = np.random.default_rng(seed)
    b = W.make_frenet_batch(B, N)
    b["kpoly"] = np.stack([rng.uniform(-2e-6, 2e-6, B), rng.uniform(-1e-4, 1e-4, B), rng.uniform(-2e-3, 2e-3, B),
                           rng.uniform(-0.05, 0.05, B)], axis=1)
    b["state"][:, 1] = rng.uniform(-1.0, 1.0, B)
    b["state"][:, 2] = rng.uniform(-0.3, 0.3, B)
    return b


@pytest.mark.parametrize("N,B,stress", [(8, 64, False), (20, 40, False), (3, 12, False), (8, 48, True), (20, 32, True), (40, 6, True)])
def test_thread_per_problem_frenet_matches_oracle(oracle, N, B, stress):
    """The Frenet-frame variant (MKZMPCPathFollowerFrenet.jl) in the thread-per-problem layout: dense s / e_y columns of the stage
    Jacobian, the 5x5 Lagrangian-Hessian block, the dense costate recursion -- against the oracle iterate for iterate, and against the
    emulated warp kernel."""
    import emu as E
    cfg = oracle.default_cfg_frenet(N)
    b = _frenet_stress_batch(B, N, 5) if stress else W.make_frenet_batch(B, N)
    o = oracle.solve_batch_frenet(cfg, b["state"], b["kpoly"], b["v_des"], b["u_prev"], want_traj=True, n_threads=4)
    e = E.solve_batch_frenet(E.kcfg_from_oracle(cfg), b["state"], b["kpoly"], b["v_des"], b["u_prev"], want_traj=True, tpp=True, slots=33)
    assert (o["status"] == e["status"]).all() and (o["iters"] == e["iters"]).all()
    ok = o["status"] == 0
    assert ok.sum() >= B - 1
    assert np.abs(o["u0"] - e["u0"])[ok].max() <= 1e-9
    assert (np.abs(o["cost"] - e["cost"])[ok] <= 1e-9 * np.maximum(1, np.abs(o["cost"][ok]))).all()
    assert np.abs(o["traj"] - e["traj"])[ok].max() <= 1e-8
    wo, we = o["traj"].copy(), o["traj"].copy()
    o2 = oracle.solve_batch_frenet(cfg, b["state"], b["kpoly"], b["v_des"], b["u_prev"], warm=wo, n_threads=4)
    e2 = E.solve_batch_frenet(E.kcfg_from_oracle(cfg), b["state"], b["kpoly"], b["v_des"], b["u_prev"], warm=we, tpp=True)
    assert (o2["status"] == e2["status"]).all() and (o2["iters"] == e2["iters"]).all()
    assert np.abs(o2["u0"] - e2["u0"])[o2["status"] == 0].max() <= 1e-9
    if N <= 20:
        w = E.solve_batch_frenet(E.kcfg_from_oracle(cfg), b["state"][:8], b["kpoly"][:8], b["v_des"][:8], b["u_prev"][:8], want_traj=True)
        assert (w["status"] == e["status"][:8]).all() and (w["iters"] == e["iters"][:8]).all() and np.abs(w["traj"] - e["traj"][:8]).max() <= 1e-9
    # MPCB200_START_ROLLOUT rolls the Frenet map out
    k1 = E.kcfg_from_oracle(cfg, start_mode=1)
    r = E.solve_batch_frenet(k1, b["state"][:8], b["kpoly"][:8], b["v_des"][:8], b["u_prev"][:8], tpp=True)
    both = (o["status"][:8] == 0) & (r["status"] == 0)
    assert both.sum() >= 6 and np.abs(o["u0"][:8] - r["u0"])[both].max() <= 1e-5


@pytest.mark.parametrize("model", ["xy", "frenet"])
def test_fuzz_classes(oracle, model):
    """One small round of every input class of tests/fuzz_layouts.py (pose errors of metres, speeds and previous commands at and
    outside their bounds, random weights, stand-still, jumping references; Frenet: also 1 - e_y K <= 0) through the oracle and
    both emulated device layouts: wherever the solves are short enough to be comparable (under 100 iterations on both sides,
    inside the model's domain) status and first move agree; infeasible inputs end Infeasible everywhere."""
    import emu as E
    import fuzz_layouts as F
    N, B = 8, 12
    rng = np.random.default_rng(7)
    seen = set()
    for kind in range(F.XY_KINDS if model == "xy" else F.FRENET_KINDS):
        if model == "xy":
            st, ref, vd, up, weights = F.perturb_xy(rng, kind, W.make_batch(B, N, b0=1000 * kind), N)
            cfg = oracle.default_cfg(N, weights=weights)
            k = E.kcfg_from_oracle(cfg)
            o = oracle.solve_batch(cfg, st, ref, vd, up, n_threads=4)
            gs = [E.solve_batch(k, st, ref, vd, up), E.solve_batch_tpp(k, st, ref, vd, up, slots=5)]
        else:
            st, kp, vd, up, weights = F.perturb_frenet(rng, kind, W.make_frenet_batch(B, N, b0=1000 * kind), N)
            cfg = oracle.default_cfg_frenet(N, weights=weights)
            k = E.kcfg_from_oracle(cfg)
            o = oracle.solve_batch_frenet(cfg, st, kp, vd, up, n_threads=4)
            gs = [E.solve_batch_frenet(k, st, kp, vd, up), E.solve_batch_frenet(k, st, kp, vd, up, tpp=True, slots=5)]
        seen.update(int(s) for s in o["status"])
        for g in gs:
            short = (g["iters"] < 100) & (o["iters"] < 100)
            if model == "frenet" and kind == 3:
                short &= (1.0 - st[:, 1] * kp[:, 3]) > 0.05    # the vehicle beyond the centre of curvature: outside the model
            sm, dm, _ = F.disagreements(g, o)
            assert not short[sm].any() and not short[dm].any(), (model, kind, sm, dm)
            assert ((g["status"] == 1) == (o["status"] == 1)).all()    # Infeasible is decided a priori, the same everywhere
    assert 0 in seen and 1 in seen
