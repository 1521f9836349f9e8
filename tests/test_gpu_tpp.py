"""GPU parity tests of the thread-per-problem layout of the solver (csrc/tpp_solver.cuh, mpc_solve_tpp_kernel), through the
C ABI: forced with mpcb200_set_large_batch_path(1) on small batches, and under its default rule on a large one.  Same
tolerances as tests/test_gpu_parity.py (BASELINE.json): same status, |du| <= 1e-5, relative cost <= 1e-6."""
import numpy as np
import pytest

from mkz_mpc_path_follower_b200 import workload as W

pytestmark = pytest.mark.gpu

U_TOL = 1e-5
COST_RTOL = 1e-6


@pytest.fixture(scope="module")
def capi():
    from mkz_mpc_path_follower_b200 import capi as c
    c.lib()
    return c


def _ocfg(oracle, solver):
    c = solver.cfg
    return oracle.default_cfg(c.N, tol=c.tol, max_iter=c.max_iter)


def _compare(g, o, min_conv=0.9, max_status_mismatch=0):
    mism = np.nonzero(g["status"] != o["status"])[0]
    assert mism.size <= max_status_mismatch, mism
    ok = (o["status"] == 0) & (g["status"] == 0)
    assert ok.mean() >= min_conv
    assert np.abs(g["u0"] - o["u0"])[ok].max() <= U_TOL
    assert (np.abs(g["cost"] - o["cost"])[ok] / np.maximum(1.0, np.abs(o["cost"][ok]))).max() <= COST_RTOL
    return ok


@pytest.mark.parametrize("N,B,paths", [(8, 300, (1,)), (20, 200, (1, 2, 3)), (3, 80, (2,)), (31, 72, (3,))])
def test_tpp_cold_and_warm_parity(capi, oracle, N, B, paths):
    s = capi.Solver(N)
    s.set_large_batch_path(1)
    b = W.make_batch(B, N, path_ids=paths)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    assert s.stats()["kernel_launches"] == 1 and B > 64
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=8)
    ok = _compare(g, o, min_conv=0.7 if N == 31 else 0.9, max_status_mismatch=3 if N == 31 else 0)
    assert np.abs(g["traj"] - o["traj"])[ok].max() <= 1e-5
    if N <= 20:
        assert (g["iters"] == o["iters"]).mean() >= 0.99
    # warm start from the solution: in/out buffer, a handful of iterations, identical counts
    wg, wo = o["traj"].copy(), o["traj"].copy()
    g2 = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], warm=wg)
    o2 = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], warm=wo, n_threads=8)
    ok2 = _compare(g2, o2, min_conv=0.7 if N == 31 else 0.9, max_status_mismatch=3 if N == 31 else 0)
    assert (g2["iters"][ok2] == o2["iters"][ok2]).mean() >= 0.99
    assert np.abs(wg - wo)[ok2].max() <= 1e-5
    # ... and the same batch through the warp-per-problem kernel
    s.set_large_batch_path(0)
    w = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    assert (w["status"] == g["status"]).sum() >= B - (3 if N == 31 else 0)
    both = (w["status"] == 0) & (g["status"] == 0)
    assert np.abs(w["u0"] - g["u0"])[both].max() <= U_TOL


@pytest.mark.parametrize("N,B", [(40, 96), (80, 80)])
def test_tpp_long_horizons(capi, oracle, N, B):
    """The thread-per-problem passes are loops over the stages: any horizon, no team size."""
    s = capi.Solver(N)
    s.set_large_batch_path(1)
    b = W.make_batch(B, N)
    w0 = W.reference_start(b, N)
    wg, wo = w0.copy(), w0.copy()
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], warm=wg)
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], warm=wo, n_threads=8)
    ok = _compare(g, o, min_conv=0.99)
    assert (g["iters"][ok] == o["iters"][ok]).all()
    assert np.abs(wg - wo)[ok].max() <= 1e-5


def test_tpp_edge_cases_records_and_restorations(capi, oracle):
    from mkz_mpc_path_follower_b200 import sharding
    N = 8
    s = capi.Solver(N)
    s.set_large_batch_path(1)
    b = W.make_batch(96, N)      # (host batches of up to 64 problems always take the packed small-batch path)
    st = b["state"].copy(); up = b["u_prev"].copy()
    st[0, 3] = 25.0; st[1, 3] = -1.0; up[2, 0] = 0.9       # infeasible by the initial speed / the previous command
    g = s.solve_batch(st, b["ref"], up, v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), st, b["ref"], b["v_des"], up, n_threads=4)
    assert (g["status"] == o["status"]).all() and (g["status"][:3] == 1).all() and (g["iters"][:3] == 0).all()
    # iteration cap -> UserLimit with the iterate returned
    s2 = capi.Solver(config=capi.default_config(N, max_iter=7))
    s2.set_large_batch_path(1)
    g = s2.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s2), b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=4)
    assert (g["status"] == 3).all() and (o["status"] == 3).all() and np.abs(g["u0"] - o["u0"]).max() <= 1e-8
    # records and restoration counts (N = 20: a few per cent of the cold starts are restored by rollout)
    N = 20
    s = capi.Solver(N)
    s.set_large_batch_path(1)
    b = W.make_batch(512, N)
    rec = s.solve_batch_records(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    a, c, stt, it, r = sharding.unpack_records_np(rec)
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=8)
    assert (stt == o["status"]).all() and (r == o["n_resto"]).all() and r.sum() > 0
    assert (s.restorations(512) == o["n_resto"]).all()
    ok = o["status"] == 0
    assert np.abs(a - o["u0"])[ok].max() <= U_TOL and (it[ok] == o["iters"][ok]).mean() >= 0.99
    # empty batch
    e = s.solve_batch(np.zeros((0, 4)), np.zeros((0, 3, N + 1)), np.zeros((0, 2)))
    assert e["u0"].shape == (0, 2)


def test_tpp_default_rule_large_batch(capi, oracle):
    """N = 8, 32,768 problems: the default rule sends the batch to the thread-per-problem kernel (lanes refill from the
    queue: 37,888 resident lanes at most, so every lane gets a problem and most are refilled at 65,536).  Every problem
    is checked against the oracle; the warp-per-problem kernel on the same batch gives the same statuses."""
    import torch
    N, B = 8, 65536
    s = capi.Solver(N)
    b = W.make_batch(B, N)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
    out = {}
    for name, mb in (("tpp", -1), ("warp", 0)):
        s.set_large_batch_path(mb)
        u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
        cost = torch.empty(B, dtype=torch.float64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        st = torch.cuda.Stream(device=dev)
        s.set_stream(st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], cost=cost, status=status, iters=iters)
            e1.record(st)
        torch.cuda.synchronize()
        out[name] = dict(u0=u0.cpu().numpy(), cost=cost.cpu().numpy(), status=status.cpu().numpy(), iters=iters.cpu().numpy(), ms=e0.elapsed_time(e1))
        s.set_stream(0)
    t, w = out["tpp"], out["warp"]
    assert (t["status"] == w["status"]).mean() >= 0.9999
    assert (t["iters"] == w["iters"]).mean() >= 0.999
    both = (t["status"] == 0) & (w["status"] == 0)
    assert (np.abs(t["u0"] - w["u0"])[both].max(axis=1) > U_TOL).sum() <= 3
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=16)
    assert (t["status"] != o["status"]).sum() <= 3
    ok = (t["status"] == 0) & (o["status"] == 0)
    assert ok.mean() >= 0.999
    assert (np.abs(t["u0"] - o["u0"])[ok].max(axis=1) > U_TOL).sum() <= 3
    assert (t["iters"][ok] == o["iters"][ok]).mean() >= 0.998
    # (no assertion on the times: the first launch of a layout allocates its device buffers inside the timed region;
    # bench.py's layouts_n8 key measures the two layouts)
    print("N=8 B=65536, first launch of each: thread-per-problem %.2f ms, warp-per-problem %.2f ms" % (t["ms"], w["ms"]))


def test_tpp_multi_device_handle(capi):
    """A handle over several GPUs with the thread-per-problem layout on every slice: bit-identical to one device (a problem's
    result does not depend on its slot or on its neighbours in the warp).  With one GPU the test is the n_devices = 1 path."""
    import torch
    N, B = 8, 3000
    b = W.make_batch(B, N)
    one = capi.Solver(N)
    one.set_large_batch_path(1)
    ref = one.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    ndev = min(torch.cuda.device_count(), 4)
    multi = capi.Solver(config=capi.default_config(N, devices=list(range(ndev))))
    multi.set_large_batch_path(1)
    out = multi.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    for k in ("u0", "cost", "status", "iters", "traj"):
        assert np.array_equal(out[k], ref[k]), k
    assert multi.stats()["kernel_launches"] == ndev
    assert np.array_equal(multi.restorations(B), one.restorations(B))


def test_tpp_nonfinite_inputs_weights_and_standstill(capi, oracle):
    """Non-finite inputs end with a status (never a hang), the same one as the oracle's; custom weights and v_des;
    the standstill start on the bound v_min (vehicle_simulator.py:31); v_des NULL."""
    N, B = 8, 96
    s = capi.Solver(N)
    s.set_large_batch_path(1)
    b = W.make_batch(B, N)
    st = b["state"].copy(); rf = b["ref"].copy(); up = b["u_prev"].copy()
    st[0, 0] = np.nan; st[1, 3] = np.inf; rf[2, 0, 3] = np.nan; up[3, 1] = np.nan
    g = s.solve_batch(st, rf, up, v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), st, rf, b["v_des"], up, n_threads=4)
    assert (g["status"] == o["status"]).all() and (g["status"][:4] != capi.OPTIMAL).all()
    assert (g["iters"][:4] == o["iters"][:4]).all()
    _compare({k: v[4:] for k, v in g.items() if v is not None}, {k: v[4:] for k, v in o.items() if k in ("u0", "cost", "status")})
    # weights and a desired speed
    w = [5.0, 7.0, 20.0, 3.0, 50.0, 500.0, 0.5, 2.0]
    s.set_cost(w)
    b2 = W.make_batch(B, N, v_des=6.0)
    g = s.solve_batch(b2["state"], b2["ref"], b2["u_prev"], v_des=b2["v_des"])
    o = oracle.solve_batch(oracle.default_cfg(N, weights=w, max_iter=s.cfg.max_iter), b2["state"], b2["ref"], b2["v_des"], b2["u_prev"], n_threads=8)
    _compare(g, o, min_conv=0.8)
    # standstill, v_des NULL
    s3 = capi.Solver(N)
    s3.set_large_batch_path(1)
    b3 = W.make_batch(B, N, path_ids=(3,))
    st = b3["state"].copy(); st[:, 3] = 0.0
    up = np.zeros((B, 2))
    g = s3.solve_batch(st, b3["ref"], up)
    o = oracle.solve_batch(_ocfg(oracle, s3), st, b3["ref"], None, up, n_threads=4)
    _compare(g, o, min_conv=0.5)


def test_tpp_closed_loop_fleet(capi, oracle):
    """mpcb200_rollout with the per-period pipeline (plant / waypoints / thread-per-problem solve over the whole fleet) against
    the fused rollout kernel on the same fleet, and a few vehicles against the oracle's closed loop."""
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N, B, T = 8, 700, 40
    trajs = [GPSRefTrajectory(mat_filename=p) for p in (1, 2, 3)]
    rng = np.random.default_rng(21)
    path_of = (np.arange(B) % 3).astype(np.int32)
    pose0 = np.stack([trajs[p].trajectory[(37 * (i + 1)) % (trajs[p].trajectory.shape[0] - 200), [4, 5, 3]] + rng.normal(scale=[0.3, 0.3, 0.03])
                      for i, p in enumerate(path_of)])
    pose0[5] = trajs[path_of[5]].trajectory[-30, [4, 5, 3]]      # runs into the stop latch
    out = {}
    for name, mb in (("warp", 0), ("tpp", 1)):
        s = capi.Solver(N)
        s.set_large_batch_path(mb)
        for i, g in enumerate(trajs):
            s.set_path(i, g.trajectory)
        out[name] = s.rollout(pose0, path_of, T)
        out[name + "_launches"] = s.stats()["kernel_launches"]
    assert out["warp_launches"] == 1 and out["tpp_launches"] == 3 * T + 1
    lw, lt = out["warp"]["log"], out["tpp"]["log"]
    assert np.array_equal(lw[:, :, 6], lt[:, :, 6])                       # statuses (and the stop latch: -1)
    assert (lw[:, :, 7] == lt[:, :, 7]).mean() >= 0.999                   # iteration counts
    # Vehicles that keep driving agree to rounding.  A vehicle behind the stop latch brakes to a standstill, where the plant is
    # discontinuous (vehicle_simulator.py:77,88-89: slip angles and lateral dynamics switch off at vx <= 1e-6, vx is clamped at
    # 0): differences of 1e-15 before that point come out as ~1e-6 in the resting pose.
    d = np.abs(lw[:, :, 0:6] - lt[:, :, 0:6]).max(axis=(0, 2))
    same = (lw[:, :, 7] == lt[:, :, 7]).all(axis=0)
    driving = (lt[:, :, 6] != -1).all(axis=0)
    print("closed loop: max |warp - tpp| %.2e over the %d vehicles that keep driving, %.2e over the %d that stop" %
          (d[driving].max(), driving.sum(), d[~driving].max() if (~driving).any() else 0.0, (~driving).sum()))
    assert same.mean() >= 0.97 and d[driving & same].max() <= 1e-8 and d.max() <= 1e-4 and driving.sum() >= B // 2
    assert np.abs(out["warp"]["final_state"] - out["tpp"]["final_state"]).max() <= 1e-4
    assert (lt[:, 5, 6] == -1).any() and (lt[-1, 5, 4:6] == [-1.0, 0.0]).all()
    # distance mode (track_with_time = false, mpc_cmd_pub.jl:99-101), a target speed, no log, non-positive target speed
    res = {}
    for name, mb in (("warp", 0), ("tpp", 1)):
        s2 = capi.Solver(N)
        s2.set_large_batch_path(mb)
        for i, g in enumerate(trajs):
            s2.set_path(i, g.trajectory)
        res[name] = (s2.rollout(pose0[:200], path_of[:200], 25, track_using_time=False, target_vel=6.5),
                     s2.rollout(pose0[:200], path_of[:200], 10, track_using_time=False, target_vel=-2.0),
                     s2.rollout(pose0[:200], path_of[:200], 10, track_using_time=False, target_vel=6.5, want_log=False))
    a_, b_ = res["warp"][0], res["tpp"][0]
    keep = (b_["log"][:, :, 6] != -1).all(axis=0)
    assert np.array_equal(a_["log"][:, :, 6:8], b_["log"][:, :, 6:8]) and np.abs(a_["log"][:, keep, 0:6] - b_["log"][:, keep, 0:6]).max() <= 1e-8
    assert np.abs(res["warp"][2]["final_state"] - res["tpp"][2]["final_state"])[keep].max() <= 1e-8
    # a non-positive target speed is a desired speed of 0 (mpc_cmd_pub.jl:58-62): the vehicles creep at ~1e-3 m/s, where the
    # plant's slip angles atan2(vy + lf wz, vx) turn differences of 1e-5 in a command into 1e-4 m: statuses, not poses
    a_, b_ = res["warp"][1], res["tpp"][1]
    assert np.array_equal(a_["log"][:, :, 6], b_["log"][:, :, 6]) and np.abs(a_["log"][:, :, 0:6] - b_["log"][:, :, 0:6]).max() <= 0.1
    cfg = _ocfg(oracle, s)
    for b in (0, 1, 2, 5, 333):
        path, keep = oracle.make_path(trajs[path_of[b]].trajectory)
        olog = oracle.closed_loop(cfg, path, pose0[b], T)
        assert np.array_equal(lt[:, b, 6], olog[:, 6]), b
        assert np.abs(lt[:, b, 0:6] - olog[:, 0:6]).max() <= (1e-8 if np.array_equal(lt[:, b, 7], olog[:, 7]) else 1e-4), b


def test_tpp_frenet_closed_loop_fleet(capi):
    """mpcb200_rollout_frenet as a per-period pipeline (plant / path ahead + the two cubic fits / thread-per-problem Frenet solve)
    against the fused kernel on the same fleet: same statuses and iteration counts, poses and commands equal to rounding."""
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N, B, T = 8, 300, 25
    trajs = [GPSRefTrajectory(mat_filename=p) for p in (1, 2, 3)]
    rng = np.random.default_rng(33)
    path_of = (np.arange(B) % 3).astype(np.int32)
    pose0 = np.stack([trajs[p].trajectory[(53 * (i + 1)) % (trajs[p].trajectory.shape[0] - 1500), [4, 5, 3]] + rng.normal(scale=[0.3, 0.3, 0.03])
                      for i, p in enumerate(path_of)])
    out = {}
    for name, mb in (("warp", 0), ("tpp", 1)):
        s = capi.FrenetSolver(N)
        s.set_large_batch_path(mb)
        for i, g in enumerate(trajs):
            s.set_path(i, g.trajectory)
        out[name] = s.rollout(pose0, path_of, T)
        out[name + "_launches"] = s.stats()["kernel_launches"]
    assert out["warp_launches"] == 1 and out["tpp_launches"] == 3 * T + 1
    lw, lt = out["warp"]["log"], out["tpp"]["log"]
    assert np.array_equal(lw[:, :, 6:8], lt[:, :, 6:8]) and (lt[:, :, 6] == 0).mean() >= 0.99
    assert np.abs(lw[:, :, 0:6] - lt[:, :, 0:6]).max() <= 1e-8
    assert np.abs(out["warp"]["final_state"] - out["tpp"]["final_state"]).max() <= 1e-8


@pytest.mark.parametrize("N", [8, 20])
def test_tpp_solve_batch_on_path(capi, oracle, N):
    """mpcb200_solve_batch_on_path through the thread-per-problem layout (on_path_ref_kernel -> mpc_solve_tpp_kernel ->
    on_path_invalid_kernel): waypoints bit-equal to the host restatement's, stop flags, the solves equal solve_batch on
    those waypoints and the warp-per-problem kernel's answers; v_des defaulting to the clamped target speed; distance mode;
    a path id outside the tables under DEVICE pointers."""
    import torch
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    s, w = capi.Solver(N), capi.Solver(N)
    s.set_large_batch_path(1)
    w.set_large_batch_path(0)
    trajs = [GPSRefTrajectory(mat_filename=p, traj_horizon=N, traj_dt=0.2) for p in (1, 2, 3)]
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
        w.set_path(i, g.trajectory)
    B = 300
    b = W.make_batch(B, N)
    path_of = (b["path"] - 1).astype(np.int32)
    g1 = s.solve_batch_on_path(b["state"], path_of, b["u_prev"], v_des=b["v_des"], want_ref=True)
    assert s.stats()["kernel_launches"] == 3
    assert np.array_equal(g1["ref"], b["ref"])
    hstop = np.zeros(B, dtype=bool)
    for p in range(3):
        m = path_of == p
        _, hstop[m] = trajs[p].get_waypoints_batch(b["state"][m, 0], b["state"][m, 1], b["state"][m, 2])
    assert np.array_equal(g1["stop"] != 0, hstop)
    g0 = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    assert np.array_equal(g1["status"], g0["status"]) and np.array_equal(g1["u0"], g0["u0"]) and np.array_equal(g1["iters"], g0["iters"])
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=8)
    _compare(g1, o, min_conv=0.95, max_status_mismatch=1)
    # without ref_out (the library's own buffer), without v_des (des_speed = max(target_vel, 0)), distance mode
    for kw in ({}, {"track_using_time": False, "target_vel": 6.0}, {"track_using_time": False, "target_vel": -3.0}):
        gt = s.solve_batch_on_path(b["state"][:128], path_of[:128], b["u_prev"][:128], **kw)
        gw = w.solve_batch_on_path(b["state"][:128], path_of[:128], b["u_prev"][:128], **kw)
        assert w.stats()["kernel_launches"] == 1
        assert np.array_equal(gt["stop"], gw["stop"])
        assert (gt["status"] == gw["status"]).mean() >= 0.98
        both = (gt["status"] == 0) & (gw["status"] == 0)
        assert both.mean() > 0.5 and np.abs(gt["u0"] - gw["u0"])[both].max() <= 1e-7
    # DEVICE pointers and a path id that was never set: status Error, zero commands, neighbours solved
    one = capi.Solver(N)
    one.set_large_batch_path(1)
    one.set_path(0, trajs[0].trajectory)
    b1 = W.make_batch(70, N, path_ids=(1,))
    dev = torch.device("cuda", 0)
    st = torch.from_numpy(b1["state"]).to(dev); up = torch.from_numpy(b1["u_prev"]).to(dev)
    pid_h = np.zeros(70, dtype=np.int32)
    pid_h[[1, 3, 4, 69]] = [5, -2, 1, 2]
    pid = torch.from_numpy(pid_h).to(dev)
    u0 = torch.full((70, 2), 7.0, dtype=torch.float64, device=dev); status = torch.full((70,), 9, dtype=torch.int32, device=dev)
    iters = torch.full((70,), 9, dtype=torch.int32, device=dev); stop = torch.full((70,), 9, dtype=torch.int32, device=dev)
    rec = torch.zeros((70, 4), dtype=torch.float64, device=dev)
    one.solve_batch_on_path_device(70, st, pid, up, u0, status=status, iters=iters, stop=stop)
    torch.cuda.synchronize()
    bad = [1, 3, 4, 69]
    good = [i for i in range(70) if i not in bad]
    assert (status.cpu().numpy()[bad] == 4).all() and (u0.cpu().numpy()[bad] == 0.0).all() and (iters.cpu().numpy()[bad] == 0).all()
    assert (stop.cpu().numpy()[bad] == 0).all()
    ref = one.solve_batch_on_path(b1["state"], np.zeros(70, dtype=np.int32), b1["u_prev"])
    assert np.array_equal(status.cpu().numpy()[good], ref["status"][good]) and np.array_equal(u0.cpu().numpy()[good], ref["u0"][good])
    del rec
