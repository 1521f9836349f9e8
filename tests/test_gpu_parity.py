"""GPU parity tests: the CUDA path, called through the C ABI (libmpc_b200.so), against the CPU
oracle on the same seeded inputs.  Tolerances are BASELINE.json's: same status, |du| <= 1e-5 on
(acc, d_f), relative cost <= 1e-6."""
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
    du = np.abs(g["u0"] - o["u0"])[ok]
    assert du.max() <= U_TOL, du.max()
    rc = np.abs(g["cost"] - o["cost"])[ok] / np.maximum(1.0, np.abs(o["cost"][ok]))
    assert rc.max() <= COST_RTOL, rc.max()
    return ok


@pytest.mark.parametrize("N,B,paths", [(8, 256, (1,)), (20, 192, (1, 2, 3)), (3, 32, (2,)), (31, 24, (3,))])
def test_cold_parity(capi, oracle, N, B, paths):
    s = capi.Solver(N)
    b = W.make_batch(B, N, path_ids=paths)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=8)
    # N = 31 is outside BASELINE.json's configs; from the all-zero start a few of its problems end in
    # "restoration needed" and rounding decides which, so one status flip is tolerated there only
    ok = _compare(g, o, min_conv=0.7 if N == 31 else 0.9, max_status_mismatch=1 if N == 31 else 0)
    assert np.abs(g["traj"] - o["traj"])[ok].max() <= 1e-5
    st = s.stats()
    assert st["kernel_launches"] == 1 and st["h2d_bytes"] > 0 and st["d2h_bytes"] > 0


@pytest.mark.parametrize("N,B", [(32, 48), (40, 256), (63, 48), (64, 48), (80, 128), (95, 32)])
def test_long_horizon_block_per_problem(capi, oracle, N, B):
    """BASELINE.json configs[4] horizons (N = 40, 80) and the team-size edges: one block of 2 or 3
    warps per problem.  Start = the reference waypoints; a zero-start slice checks status parity."""
    s = capi.Solver(N)
    b = W.make_batch(B, N)
    ocfg = _ocfg(oracle, s)
    w0 = W.reference_start(b, N)
    wo, wg = w0.copy(), w0.copy()
    o = oracle.solve_batch(ocfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=wo, n_threads=8)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], warm=wg)
    ok = _compare(g, o, min_conv=0.95)
    assert (g["iters"][ok] == o["iters"][ok]).mean() > 0.98
    assert np.abs(wg - wo)[ok].max() <= 1e-5
    assert s.stats()["kernel_launches"] == 1
    # the all-zero start of the spec: at these horizons many problems run 100-200 iterations from a
    # point hundreds of metres away, and rounding decides on which side of the cap (or of the
    # line-search failure) a few of them end; the short runs must agree exactly
    nz = min(B, 24)
    g0 = s.solve_batch(b["state"][:nz], b["ref"][:nz], b["u_prev"][:nz], v_des=b["v_des"][:nz])
    o0 = oracle.solve_batch(ocfg, b["state"][:nz], b["ref"][:nz], b["v_des"][:nz], b["u_prev"][:nz], n_threads=8)
    short = o0["iters"] < 100
    assert (g0["status"][short] == o0["status"][short]).all()
    assert (g0["iters"][short] == o0["iters"][short]).all()
    both = (g0["status"] == 0) & (o0["status"] == 0)
    if both.any():
        assert np.abs(g0["u0"] - o0["u0"])[both].max() <= U_TOL


@pytest.mark.parametrize("N,B", [(8, 128), (20, 96)])
def test_warm_parity_and_inout_buffer(capi, oracle, N, B):
    s = capi.Solver(N)
    b = W.make_batch(B, N)
    ocfg = _ocfg(oracle, s)
    o1 = oracle.solve_batch(ocfg, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=8)
    # perturb the state a little: "next control step" from the previous solution
    rng = np.random.default_rng(5)
    st2 = b["state"] + rng.normal(scale=[0.05, 0.05, 0.005, 0.05], size=(B, 4))
    st2[:, 3] = np.clip(st2[:, 3], 0, 20)
    warm_o = o1["traj"].copy(); warm_g = o1["traj"].copy()
    o2 = oracle.solve_batch(ocfg, st2, b["ref"], b["v_des"], b["u_prev"], warm=warm_o, n_threads=8)
    g2 = s.solve_batch(st2, b["ref"], b["u_prev"], v_des=b["v_des"], warm=warm_g)
    ok = _compare(g2, o2)
    assert np.abs(warm_g - warm_o)[ok].max() <= 1e-5      # warm is written back with the solution
    assert g2["iters"][ok].mean() < 20


def test_weights_and_vdes(capi, oracle):
    N, B = 8, 64
    w = [5.0, 7.0, 20.0, 3.0, 50.0, 500.0, 0.5, 2.0]
    s = capi.Solver(N); s.set_cost(w)
    b = W.make_batch(B, N, v_des=6.0)
    ocfg = oracle.default_cfg(N, weights=w, max_iter=s.cfg.max_iter)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    o = oracle.solve_batch(ocfg, b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=8)
    _compare(g, o, min_conv=0.8)


def test_edge_cases(capi, oracle):
    N = 8
    s = capi.Solver(N)
    b = W.make_batch(4, N)
    # empty batch
    g = s.solve_batch(b["state"][:0], b["ref"][:0], b["u_prev"][:0])
    assert g["u0"].shape == (0, 2)
    # batch of one, v_des NULL
    g = s.solve_batch(b["state"][:1], b["ref"][:1], b["u_prev"][:1])
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"][:1], b["ref"][:1], None, b["u_prev"][:1])
    _compare(g, o, min_conv=0.0)
    # infeasible inputs: v0 < v_min, previous steering outside box + rate step
    st = b["state"].copy(); st[0, 3] = -0.5
    up = b["u_prev"].copy(); up[1, 0] = 0.7
    g = s.solve_batch(st, b["ref"], up, v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), st, b["ref"], b["v_des"], up)
    assert g["status"][0] == capi.INFEASIBLE and g["status"][1] == capi.INFEASIBLE
    assert (g["status"] == o["status"]).all()
    # iteration cap -> UserLimit
    s2 = capi.Solver(N, max_iter=3)
    g = s2.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    assert (g["status"] == capi.USERLIMIT).all() and (g["iters"] == 3).all()
    # non-finite inputs: every solve terminates with a status (never a hang), the same as the oracle's
    st = b["state"].copy(); rf = b["ref"].copy(); up = b["u_prev"].copy()
    st[0, 0] = np.nan; st[1, 3] = np.inf; rf[2, 0, 3] = np.nan; up[3, 1] = np.nan
    g = s.solve_batch(st, rf, up, v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), st, rf, b["v_des"], up)
    assert (g["status"] == o["status"]).all() and (g["status"] != capi.OPTIMAL).all()
    assert (g["iters"] == o["iters"]).all()
    # argument validation
    with pytest.raises(ValueError):
        s.solve_batch(b["state"], b["ref"][:, :, :-1], b["u_prev"])
    with pytest.raises(capi.MpcB200Error):
        capi.Solver(96)   # horizons up to 95 (three warps per problem)
    with pytest.raises(capi.MpcB200Error):
        capi.Solver(2)
    with pytest.raises(capi.MpcB200Error):
        s.set_cost([-1.0] * 8)


def test_standstill_start(capi, oracle):
    """The simulator starts at v = 0 exactly on the bound v_min (vehicle_simulator.py:31)."""
    N = 8
    s = capi.Solver(N)
    b = W.make_batch(16, N, path_ids=(3,))
    st = b["state"].copy(); st[:, 3] = 0.0
    up = np.zeros((16, 2))
    g = s.solve_batch(st, b["ref"], up, v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), st, b["ref"], b["v_des"], up, n_threads=4)
    _compare(g, o, min_conv=0.5)


def _feasibility(cfg, b, g, N):
    """size-independent property: a point returned as Optimal satisfies the reference NLP's
    constraints (dynamics of MKZMPCPathFollower.jl:115-123, bounds, rate rows)."""
    t = g["traj"]
    x, y, v, psi = (t[:, i * (N + 1):(i + 1) * (N + 1)] for i in range(4))
    df = t[:, 4 * (N + 1):4 * (N + 1) + N]; acc = t[:, 4 * (N + 1) + N:]
    r = cfg.L_b / (cfg.L_a + cfg.L_b)
    bta = np.arctan(r * np.tan(df))
    res = np.stack((x[:, 1:] - (x[:, :-1] + cfg.dt * v[:, :-1] * np.cos(psi[:, :-1] + bta)),
                    y[:, 1:] - (y[:, :-1] + cfg.dt * v[:, :-1] * np.sin(psi[:, :-1] + bta)),
                    psi[:, 1:] - (psi[:, :-1] + cfg.dt * v[:, :-1] / cfg.L_b * np.sin(bta)),
                    v[:, 1:] - (v[:, :-1] + cfg.dt * acc)))
    init = np.abs(np.stack((x[:, 0], y[:, 0], psi[:, 0], v[:, 0]), axis=1) - b["state"]).max(axis=1)
    dyn = np.abs(res).max(axis=(0, 2))
    bnd = np.maximum.reduce([(-v).max(axis=1), (v - cfg.v_max).max(axis=1), (np.abs(acc) - cfg.a_max).max(axis=1),
                             (np.abs(df) - cfg.steer_max).max(axis=1)])
    first = np.maximum(np.abs(df[:, 0] - b["u_prev"][:, 0]) - cfg.steer_dmax * cfg.dt_control,
                       np.abs(acc[:, 0] - b["u_prev"][:, 1]) - cfg.a_dmax * cfg.dt_control)
    later = np.maximum((np.abs(np.diff(df, axis=1))[:, 1:] - cfg.steer_dmax * cfg.dt).max(axis=1),
                       (np.abs(np.diff(acc, axis=1))[:, 1:] - cfg.a_dmax * cfg.dt).max(axis=1))
    return np.maximum.reduce([init, dyn, bnd, first, later])


@pytest.mark.parametrize("N,B,paths", [(8, 4096, (1,)), (20, 65536, (1, 2, 3))])
def test_full_size_properties(capi, oracle, N, B, paths):
    """BASELINE.json configs[1] and configs[2] at full size: feasibility of every Optimal point,
    slice reproducibility (any contiguous slice gives bit-identical results), idempotence of a
    warm re-solve, and oracle parity on a random subset."""
    s = capi.Solver(N)
    b = W.make_batch(B, N, path_ids=paths)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    ok = g["status"] == 0
    assert ok.mean() > 0.999    # the ~2 % whose line search fails from the all-zero start are restored by rollout
    viol = _feasibility(s.cfg, b, g, N)
    assert viol[ok].max() <= 2e-8
    # slice reproducibility
    lo, hi = B // 3, B // 3 + 500
    bs = W.make_batch(hi - lo, N, path_ids=paths, b0=lo)
    assert np.array_equal(bs["state"], b["state"][lo:hi])
    g2 = s.solve_batch(bs["state"], bs["ref"], bs["u_prev"], v_des=bs["v_des"])
    assert np.array_equal(g2["u0"], g["u0"][lo:hi]) and np.array_equal(g2["status"], g["status"][lo:hi])
    assert np.array_equal(g2["iters"], g["iters"][lo:hi])
    # idempotence
    warm = g["traj"].copy()
    g3 = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], warm=warm)
    both = ok & (g3["status"] == 0)
    assert both.sum() >= 0.99 * ok.sum()
    # two DIFFERENT iterate paths to the same KKT point agree only to the accuracy of Ipopt's stopping
    # rule (tol = 1e-8 on the gradient-scaled problem): ~1e-6 typically, ~1e-4 worst case in 64K
    d3 = np.abs(g3["u0"] - g["u0"])[both]
    assert np.median(d3) <= 1e-6 and np.quantile(d3, 0.999) <= 2e-4 and d3.max() <= 2e-3, (np.median(d3), d3.max())
    # oracle on a subset
    idx = np.random.default_rng(11).choice(B, size=384, replace=False)
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"][idx], b["ref"][idx], b["v_des"][idx], b["u_prev"][idx], n_threads=8)
    _compare({k: g[k][idx] for k in ("u0", "cost", "status")}, o)


@pytest.mark.parametrize("N,B", [(8, 512), (20, 512), (40, 64)])
def test_rollout_start_mode(capi, oracle, N, B):
    """MPCB200_START_ROLLOUT (opt-in, not a reference behaviour) against the oracle started from the same
    rolled-out point; the solves need a fraction of the all-zero start's iterations."""
    s = capi.Solver(N, start_mode=capi.START_ROLLOUT)
    b = W.make_batch(B, N)
    ocfg = _ocfg(oracle, s)
    w = oracle.rollout_start(ocfg, b["state"], b["u_prev"])
    o = oracle.solve_batch(ocfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=w.copy(), n_threads=8)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    ok = _compare(g, o, min_conv=0.99)
    assert (g["iters"][ok] == o["iters"][ok]).mean() > 0.98 and g["iters"].mean() < 15


@pytest.mark.parametrize("N", [8, 20, 40])
def test_solve_batch_on_path(capi, oracle, N):
    """mpcb200_solve_batch_on_path: waypoints generated on the device (ref_gps_traj.py:131-218) must equal the host
    restatement's (which is pinned bit-exactly to the reference's own output by tests/golden) and the solves must
    equal solve_batch on those waypoints; plus distance mode, the stop flag and argument checks."""
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    s = capi.Solver(N)
    trajs = [GPSRefTrajectory(mat_filename=p, traj_horizon=N, traj_dt=0.2) for p in (1, 2, 3)]
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
    B = 300
    b = W.make_batch(B, N)                      # paths 1-3 round-robin -> ids 0..2
    path_of = (b["path"] - 1).astype(np.int32)
    g1 = s.solve_batch_on_path(b["state"], path_of, b["u_prev"], v_des=b["v_des"], want_ref=True)
    st = s.stats()
    assert st["kernel_launches"] == 1 and st["h2d_bytes"] == B * (4 * 8 + 4 + 2 * 8 + 8)
    # time mode, heading unwrap included: every operation of the generator is rounded like numpy's, so the waypoints are
    # the host restatement's (which tests/golden pins to the reference's own output) bit for bit
    assert np.array_equal(g1["ref"], b["ref"])
    hstop = np.zeros(B, dtype=bool)
    for p in range(3):
        m = path_of == p
        _, hstop[m] = trajs[p].get_waypoints_batch(b["state"][m, 0], b["state"][m, 1], b["state"][m, 2])
    assert np.array_equal(g1["stop"] != 0, hstop)
    g0 = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    assert (g1["status"] == g0["status"]).all()
    ok = g0["status"] == 0
    assert np.array_equal(g1["u0"], g0["u0"]) and np.array_equal(g1["iters"], g0["iters"])   # same waypoints, same solve
    # distance mode (track_using_time = False): waypoints at s_closest + (h+1) dt v_target
    vt = 6.0
    refd = np.empty((B, 3, N + 1))
    for p in range(3):
        m = path_of == p
        refd[m], _ = trajs[p].get_waypoints_batch(b["state"][m, 0], b["state"][m, 1], b["state"][m, 2], v_target=vt)
    g2 = s.solve_batch_on_path(b["state"], path_of, b["u_prev"], track_using_time=False, target_vel=vt, want_ref=True)
    assert np.array_equal(g2["ref"], refd)
    if N <= 31:
        o2 = oracle.solve_batch(_ocfg(oracle, s), b["state"][:64], refd[:64], np.full(64, vt), b["u_prev"][:64], n_threads=8)
        _compare({k: g2[k][:64] for k in ("u0", "cost", "status")}, o2, min_conv=0.5)
    # a non-positive target_vel is des_speed = 0 (mpc_cmd_pub.jl:58-62): N + 1 copies of the nearest sample
    g4 = s.solve_batch_on_path(b["state"][:8], path_of[:8], b["u_prev"][:8], track_using_time=False, target_vel=-3.0, want_ref=True)
    for p in range(3):
        m = path_of[:8] == p
        r0, _ = trajs[p].get_waypoints_batch(b["state"][:8][m, 0], b["state"][:8][m, 1], b["state"][:8][m, 2], v_target=0.0)
        assert np.array_equal(g4["ref"][m], r0)
    # the end of the path raises stop_cmd
    g = trajs[0]
    j = g.trajectory.shape[0] - 50
    stt = np.array([[g.trajectory[j, 4], g.trajectory[j, 5], g.trajectory[j, 3], 5.0]])
    g3 = s.solve_batch_on_path(stt, np.array([0]), np.zeros((1, 2)))
    assert g3["stop"][0] == 1
    with pytest.raises(capi.MpcB200Error):
        capi.Solver(N).solve_batch_on_path(stt, np.array([0]), np.zeros((1, 2)))     # path not set
    with pytest.raises(capi.MpcB200Error):
        s.solve_batch_on_path(stt, np.array([3]), np.zeros((1, 2)))                  # host pointers: a bad id is refused


def test_on_path_device_pointers_bad_path_id(capi):
    """With DEVICE pointers the host cannot inspect path_of: a problem whose id is outside the tables (or names a table
    that was never set) comes back with status Error and zero commands; its neighbours are solved."""
    import torch
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N = 8
    s = capi.Solver(N)
    g = GPSRefTrajectory(mat_filename=1, traj_horizon=N, traj_dt=0.2)
    s.set_path(0, g.trajectory)                       # only table 0 exists
    b = W.make_batch(6, N, path_ids=(1,))
    dev = torch.device("cuda", 0)
    st = torch.from_numpy(b["state"]).to(dev); up = torch.from_numpy(b["u_prev"]).to(dev)
    pid = torch.tensor([0, 5, 0, -2, 1, 0], dtype=torch.int32, device=dev)
    u0 = torch.full((6, 2), 7.0, dtype=torch.float64, device=dev); status = torch.full((6,), 9, dtype=torch.int32, device=dev)
    iters = torch.full((6,), 9, dtype=torch.int32, device=dev)
    s.solve_batch_on_path_device(6, st, pid, up, u0, status=status, iters=iters)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0, 4, 0, 4, 4, 0]
    bad = [1, 3, 4]
    assert (u0.cpu().numpy()[bad] == 0.0).all() and iters.cpu().numpy()[bad].tolist() == [0, 0, 0]
    ref = s.solve_batch_on_path(b["state"], np.zeros(6, dtype=np.int32), b["u_prev"])
    good = [0, 2, 5]
    assert np.array_equal(u0.cpu().numpy()[good], ref["u0"][good])


def test_records_restorations_and_graph_cache(capi, oracle):
    """mpcb200_solve_batch_records: the kernel writes the packed 32-byte record {acc, df, cost, status | restorations << 8,
    iters} itself; mpcb200_get_restorations reports which problems went through the restoration by rollout.  Then the
    small-batch graph cache across set_cost / set_stream / other shapes: every replay uses the current weights."""
    from mkz_mpc_path_follower_b200 import sharding
    N, B = 20, 4096
    s = capi.Solver(N)
    b = W.make_batch(B, N)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    resto = s.restorations(B)
    rec = s.solve_batch_records(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    u0, cost, status, iters, nres = sharding.unpack_records_np(rec)
    assert np.array_equal(u0, g["u0"]) and np.array_equal(cost, g["cost"])
    assert np.array_equal(status, g["status"]) and np.array_equal(iters, g["iters"]) and np.array_equal(nres, resto)
    assert 0 < (resto > 0).sum() < 0.06 * B and resto.max() <= 3           # a few per cent of the cold starts, at most 3 each
    # the oracle restores the same problems (same rule, same points)
    n = 512
    cfg = _ocfg(oracle, s)
    on = oracle.solve_batch(cfg, b["state"][:n], b["ref"][:n], b["v_des"][:n], b["u_prev"][:n], n_threads=8)["n_resto"]
    assert np.array_equal(on, resto[:n])
    with pytest.raises(capi.MpcB200Error):
        s.restorations(B + 1)
    # small batches (graph replay) interleaved with weight / stream / shape changes
    s8 = capi.Solver(8)
    b8 = W.make_batch(4, 8)
    w1 = [9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0]; w2 = [1.0, 30.0, 5.0, 0.2, 10.0, 100.0, 0.1, 0.1]
    def solve(k):
        return s8.solve_batch(b8["state"][:k], b8["ref"][:k], b8["u_prev"][:k], v_des=b8["v_des"][:k])["u0"]
    a1 = solve(1); a4 = solve(4)
    s8.set_cost(w2); c1 = solve(1); c4 = solve(4)
    assert not np.array_equal(a1, c1) and np.array_equal(c1, c4[:1])
    s8.set_cost(w1)
    assert np.array_equal(solve(1), a1) and np.array_equal(solve(4), a4)
    import torch
    st = torch.cuda.Stream()
    s8.set_stream(st.cuda_stream); assert np.array_equal(solve(4), a4)
    s8.set_cost(w2); assert np.array_equal(solve(1), c1)
    s8.set_stream(None); assert np.array_equal(solve(4), c4)
    fresh = capi.Solver(8); fresh.set_cost(w2)
    assert np.array_equal(fresh.solve_batch(b8["state"], b8["ref"], b8["u_prev"], v_des=b8["v_des"])["u0"], c4)


def test_multi_device_handle(capi):
    """n_devices > 1 inside libmpc_b200.so: HOST-pointer batches and rollouts are cut into contiguous slices
    over the devices and come back bit-identical to one device.  On a one-GPU box the two 'devices' cannot be distinct:
    the duplicate is refused, and the test is the n_devices = 1 path."""
    import torch
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N, B = 8, 2001
    b = W.make_batch(B, N)
    one = capi.Solver(N)
    ref = one.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    with pytest.raises(capi.MpcB200Error):
        capi.Solver(config=capi.default_config(N, devices=[0, 0]))
    with pytest.raises(capi.MpcB200Error):
        capi.Solver(config=capi.default_config(N, devices=[0, torch.cuda.device_count()]))
    ndev = min(torch.cuda.device_count(), 4)
    multi = capi.Solver(config=capi.default_config(N, devices=list(range(ndev))))
    out = multi.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    for k in ("u0", "cost", "status", "iters", "traj"):
        assert np.array_equal(out[k], ref[k]), k
    assert multi.stats()["kernel_launches"] == ndev
    assert np.array_equal(multi.restorations(B), (one.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"]), one.restorations(B))[1])
    trajs = [GPSRefTrajectory(mat_filename=p) for p in (1, 2, 3)]
    for i, g in enumerate(trajs):
        one.set_path(i, g.trajectory); multi.set_path(i, g.trajectory)
    rng = np.random.default_rng(11)
    nv, T = 37, 12
    path_of = (np.arange(nv) % 3).astype(np.int32)
    pose0 = np.stack([trajs[p].trajectory[100 * (i + 1), [4, 5, 3]] + rng.normal(scale=[0.3, 0.3, 0.03]) for i, p in enumerate(path_of)])
    r1 = one.rollout(pose0, path_of, T); rm = multi.rollout(pose0, path_of, T)
    assert np.array_equal(r1["log"], rm["log"]) and np.array_equal(r1["final_state"], rm["final_state"])
    o1 = one.solve_batch_on_path(b["state"][:500], (b["path"][:500] - 1).astype(np.int32), b["u_prev"][:500], want_ref=True)
    om = multi.solve_batch_on_path(b["state"][:500], (b["path"][:500] - 1).astype(np.int32), b["u_prev"][:500], want_ref=True)
    for k in ("u0", "status", "iters", "ref", "stop"):
        assert np.array_equal(o1[k], om[k]), k


@pytest.mark.parametrize("layout", ["warp", "thread"])
@pytest.mark.parametrize("N,B,start", [(8, 48, "zero"), (20, 48, "zero"), (40, 12, "ref")])
def test_kkt_of_cuda_solutions(capi, oracle, N, B, start, layout):
    """Intrinsic check that does not involve the oracle's interior-point code: every Optimal point the
    CUDA solver returns satisfies the KKT conditions of the UNSCALED reference NLP (stationarity with
    least-squares multipliers on the active set, primal feasibility, multiplier signs).  The N = 20
    batch contains problems that go through the restoration by rollout."""
    import ctypes as C
    from test_oracle_solver import _kkt_residual, _p
    s = capi.Solver(N)
    s.set_large_batch_path(0 if layout == "warp" else 1)     # both device layouts of the solver (the second: csrc/tpp_solver.cuh)
    if layout == "thread":
        B += 64                                              # (host batches of up to 64 problems always take the small-batch path)
    b = W.make_batch(B, N, b0=30 if N == 20 else 0)     # problems 38, 62 of the stream need a restoration
    warm = W.reference_start(b, N) if start == "ref" else None
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], warm=warm, want_traj=True)
    assert (g["status"] == 0).mean() >= 0.95
    ocfg = _ocfg(oracle, s)
    for j in np.nonzero(g["status"] == 0)[0]:
        z = np.empty(6 * N + 4)
        oracle.lib().mpc_oracle_traj_to_z(C.byref(ocfg), _p(np.ascontiguousarray(g["traj"][j])), _p(z))
        stat, prim, sign = _kkt_residual(oracle, ocfg, (b["state"][j], b["ref"][j], 1.0, b["u_prev"][j]), z)
        assert prim <= 1e-8 * max(1.0, np.abs(z).max()), (j, prim)
        assert stat <= 1e-4, (j, stat)       # unscaled; Ipopt's tol 1e-8 acts on the scaled problem
        assert sign <= 1e-6, (j, sign)


def test_cuda_solution_is_what_scipy_slsqp_finds(capi, oracle):
    """Independent solver (scipy SLSQP, analytic derivatives of the reference NLP), started near the CUDA solver's
    point: it must stay there.  Does not involve the oracle's interior-point code."""
    import ctypes as C
    from scipy.optimize import minimize
    from test_oracle_solver import _nlp, _p
    N = 8
    s = capi.Solver(N)
    ocfg = _ocfg(oracle, s)
    b = W.make_batch(6, N)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    rng = np.random.default_rng(0)
    checked = 0
    for j in np.nonzero(g["status"] == 0)[0]:
        f, gr, c, d, jac, lo, hi, dlim = _nlp(oracle, ocfg, b["state"][j], b["ref"][j], 1.0, b["u_prev"][j])
        z0 = np.empty(6 * N + 4)
        oracle.lib().mpc_oracle_traj_to_z(C.byref(ocfg), _p(np.ascontiguousarray(g["traj"][j])), _p(z0))
        cons = [{"type": "eq", "fun": c, "jac": lambda z: jac(z)[0]},
                {"type": "ineq", "fun": lambda z: dlim - d(z), "jac": lambda z: -jac(z)[1]},
                {"type": "ineq", "fun": lambda z: dlim + d(z), "jac": lambda z: jac(z)[1]}]
        res = minimize(f, z0 + 1e-2 * rng.normal(size=z0.size), jac=gr, bounds=list(zip(lo, hi)), constraints=cons,
                       method="SLSQP", options={"ftol": 1e-15, "maxiter": 500})
        assert np.abs(c(res.x)).max() < 1e-7
        assert abs(res.fun - g["cost"][j]) <= 1e-6 * max(1.0, abs(g["cost"][j])), (j, res.fun, g["cost"][j])
        assert np.abs(res.x[4:6] - g["u0"][j]).max() < 1e-4, (j, res.x[4:6], g["u0"][j])
        checked += 1
    assert checked >= 4


def test_large_batch_oracle_parity(capi, oracle):
    """16,384 problems of configs[2] (a quarter of the bench batch) against the oracle, every one of them.  The
    two implementations round differently (Riccati vs dense LDL^T), and from the all-zero start a handful of
    problems run 70-200 iterations through nonconvex territory where that matters: measured on this slice,
    15 iteration counts differ, one problem converges on the GPU at iteration 153 while the oracle reaches
    the cap, and ONE pair ends in two different local minima (|du| = 0.1).  Everything else agrees to 2e-7."""
    N, B = 20, 16384
    s = capi.Solver(N)
    b = W.make_batch(B, N, b0=65536)     # a slice the other tests do not touch
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=16)
    assert (g["status"] != o["status"]).sum() <= 3
    ok = (g["status"] == 0) & (o["status"] == 0)
    assert ok.mean() >= 0.999
    assert (g["iters"][ok] == o["iters"][ok]).mean() >= 0.998
    du = np.abs(g["u0"] - o["u0"])[ok].max(axis=1)
    assert (du > U_TOL).sum() <= 3 and np.quantile(du, 0.999) <= 1e-7
    short = ok & (o["iters"] <= 60)      # the bulk: no chaos yet, bit-for-bit the same path
    assert (g["iters"][short] == o["iters"][short]).all() and np.abs(g["u0"] - o["u0"])[short].max() <= 1e-8


def test_julia_module_mirror(capi, oracle):
    """The six-function API of MKZMPCPathFollower.jl:132-207, batch of one, one control step."""
    from mkz_mpc_path_follower_b200.mpc_path_follower import MKZMPCPathFollower
    kmpc = MKZMPCPathFollower(N=8)
    assert kmpc.N == 8 and kmpc.dt == 0.2
    kmpc.update_cost(9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0)
    b = W.make_batch(1, 8, path_ids=(3,))
    x, y, psi, v = b["state"][0]
    kmpc.update_init_cond(x, y, psi, v)
    kmpc.update_reference(b["ref"][0, 0], b["ref"][0, 1], b["ref"][0, 2], 1.0)
    kmpc.update_current_input(b["u_prev"][0, 0], b["u_prev"][0, 1])
    a_opt, df_opt, is_opt = kmpc.solve_model()
    res = kmpc.get_solver_results()
    assert len(res) == 9 and res[0].shape == (9,) and res[7].shape == (8,)
    assert is_opt in ("Optimal", "Error", "UserLimit")
    if is_opt == "Optimal":
        assert res[8][0] == a_opt and res[7][0] == df_opt       # acc_opt[1], d_f_opt[1]
        assert abs(res[0][0] - x) < 1e-7 and abs(res[2][0] - v) < 1e-7   # x_mpc[1], v_mpc[1]: order x,y,v,psi
    with pytest.raises(TypeError):
        kmpc.update_reference(np.zeros(3), np.zeros(9), np.zeros(9), 1.0)


def test_closed_loop_reference_workload(capi, oracle):
    """BASELINE.json configs[0]: the launch file's closed-loop lane keep (path3, X0=0, Y0=3, Psi0=-1.5,
    launch/sim_path_follow.launch:13,23-25), N=8, time-mode references, warm-started; first 12 s against
    the oracle's closed loop, plus path1 from its own start."""
    from mkz_mpc_path_follower_b200 import closed_loop
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    T = 120
    for pid, pose in ((3, (0.0, 3.0, -1.5)), (1, None)):
        g = GPSRefTrajectory(mat_filename=pid)
        if pose is None:
            pose = (g.trajectory[0, 4] + 0.5, g.trajectory[0, 5] - 0.5, g.trajectory[0, 3] + 0.05)
        out = closed_loop.run([pid], [pose], T, N=8)
        path, keep = oracle.make_path(g.trajectory)
        olog = oracle.closed_loop(oracle.default_cfg(8), path, pose, T, track_using_time=True, target_vel=1.0, warm_start=True)
        glog = out["log"][:, 0, :]
        assert np.array_equal(glog[:, 6], olog[:, 6])                      # same status every step
        assert np.abs(glog[:, 4:6] - olog[:, 4:6]).max() <= 1e-5           # same published commands
        assert np.abs(glog[:, 0:4] - olog[:, 0:4]).max() <= 1e-4           # same closed-loop trajectory
        assert (glog[:, 6] == 0).mean() > 0.95
        err = closed_loop.path_errors(out["log"][20:], g.trajectory)
        assert err.max() < 1.5                                             # it follows the path


def test_device_rollout_matches_oracle_closed_loop(capi, oracle):
    """mpcb200_rollout (plant + reference generation + warm-started solves in one persistent kernel) against the
    oracle's closed loop, vehicle by vehicle: same status and iteration count every step, same commands."""
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    s = capi.Solver(8)
    trajs = [GPSRefTrajectory(mat_filename=p) for p in (1, 2, 3)]
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
    rng = np.random.default_rng(4)
    poses, path_of = [(0.0, 3.0, -1.5)], [2]           # the launch file's start on path3
    for i, g in enumerate(trajs):
        for j in (0, 1500, 4000):
            poses.append((g.trajectory[j, 4] + rng.normal(scale=0.4), g.trajectory[j, 5] + rng.normal(scale=0.4),
                          g.trajectory[j, 3] + rng.normal(scale=0.05)))
            path_of.append(i)
    T = 80
    out = s.rollout(np.array(poses), np.array(path_of), T)
    assert s.stats()["kernel_launches"] == 1          # (the module-load solve of the first call is not counted)
    for b, (pose, pid) in enumerate(zip(poses, path_of)):
        path, keep = oracle.make_path(trajs[pid].trajectory)
        olog = oracle.closed_loop(oracle.default_cfg(8), path, pose, T)
        glog = out["log"][:, b, :]
        assert np.array_equal(glog[:, 6], olog[:, 6]), b
        assert np.abs(glog[:, 4:6] - olog[:, 4:6]).max() <= 1e-5, b
        assert np.abs(glog[:, 0:4] - olog[:, 0:4]).max() <= 1e-4, b
    # the end-of-path latch: a vehicle started 3 s before the end of path1 brakes with (-1, 0)
    g = trajs[0]
    j = g.trajectory.shape[0] - 300
    out = s.rollout(np.array([[g.trajectory[j, 4], g.trajectory[j, 5], g.trajectory[j, 3]]]), np.array([0]), 60)
    assert (out["log"][-1, 0, 4:7] == np.array([-1.0, 0.0, -1.0])).all()
    with pytest.raises(capi.MpcB200Error):
        capi.Solver(8).rollout(np.zeros((1, 3)), np.array([0]), 5)      # path not set


def test_device_rollout_long_horizon(capi, oracle):
    """Closed-loop rollouts at N = 40 (one block of two warps per vehicle; the reference generator takes any
    traj_horizon, ref_gps_traj.py:60) against the oracle's closed loop."""
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N, T = 40, 12
    s = capi.Solver(N)
    g = GPSRefTrajectory(mat_filename=2, traj_horizon=N, traj_dt=0.2)
    for i in range(3):
        s.set_path(i, g.trajectory)
    rng = np.random.default_rng(6)
    poses = np.array([g.trajectory[j, [4, 5, 3]] + rng.normal(scale=[0.2, 0.2, 0.02]) for j in (300, 1800, 3300, 4100, 5000)])
    out = s.rollout(poses, np.zeros(5, dtype=np.int32), T)
    path, keep = oracle.make_path(g.trajectory)
    cfg = _ocfg(oracle, s)
    for b in range(poses.shape[0]):
        olog = oracle.closed_loop(cfg, path, poses[b], T)
        glog = out["log"][:, b, :]
        assert np.array_equal(glog[:, 6], olog[:, 6]), b
        assert np.abs(glog[:, 4:6] - olog[:, 4:6]).max() <= 1e-5, b
        assert np.abs(glog[:, 0:4] - olog[:, 0:4]).max() <= 1e-4, b


def test_monte_carlo_rollout_properties(capi):
    """BASELINE.json configs[3] scaled to one GPU-minute: 2,048 vehicles x 100 steps over paths 1-3; every step's solve
    is Optimal once the vehicles are on the path, and the fleet tracks its paths."""
    from mkz_mpc_path_follower_b200 import closed_loop
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    s = capi.Solver(8)
    trajs = [GPSRefTrajectory(mat_filename=p) for p in (1, 2, 3)]
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
    B, T = 2048, 100
    rng = np.random.default_rng(9)
    path_of = (np.arange(B) % 3).astype(np.int32)
    pose0 = np.stack([(trajs[p].trajectory[0, 4] + rng.normal(scale=0.3), trajs[p].trajectory[0, 5] + rng.normal(scale=0.3),
                       trajs[p].trajectory[0, 3] + rng.normal(scale=0.05)) for p in path_of])
    out = s.rollout(pose0, path_of, T)
    st = out["log"][:, :, 6]
    assert (st == 0).mean() > 0.97
    for p in range(3):
        err = closed_loop.path_errors(out["log"][30:, path_of == p, :], trajs[p].trajectory)
        assert np.quantile(err, 0.99) < 1.0


# ---- Frenet-frame variant (scripts/mpc_utils/MKZMPCPathFollowerFrenet.jl) ----
def _frenet_stress(b, seed):
    rng = np.random.default_rng(seed)
    B = b["state"].shape[0]
    b["kpoly"] = np.stack([rng.uniform(-2e-6, 2e-6, B), rng.uniform(-1e-4, 1e-4, B), rng.uniform(-2e-3, 2e-3, B),
                           rng.uniform(-0.05, 0.05, B)], axis=1)
    b["state"][:, 1] = rng.uniform(-1.0, 1.0, B)
    b["state"][:, 2] = rng.uniform(-0.3, 0.3, B)
    return b


@pytest.mark.parametrize("N,B,stress", [(8, 512, False), (20, 384, False), (3, 32, False), (31, 32, False), (8, 256, True), (20, 256, True),
                                            (32, 48, False), (40, 128, True), (64, 32, False), (80, 64, True), (95, 16, False)])
@pytest.mark.parametrize("layout", ["warp", "thread"])
def test_frenet_cold_and_warm_parity(capi, oracle, N, B, stress, layout):
    s = capi.FrenetSolver(N)
    s.set_large_batch_path(0 if layout == "warp" else 1)     # both device layouts (the second: TppSolverT<1>, csrc/tpp_solver.cuh)
    if layout == "thread":
        B += 64                                              # (host batches of up to 64 problems always take the small-batch path)
    b = W.make_frenet_batch(B, N)
    if stress:
        b = _frenet_stress(b, 17)
    ocfg = oracle.default_cfg_frenet(N, tol=s.cfg.tol, max_iter=s.cfg.max_iter)
    g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    o = oracle.solve_batch_frenet(ocfg, b["state"], b["kpoly"], b["v_des"], b["u_prev"], want_traj=True, n_threads=8)
    ok = _compare(g, o, min_conv=0.95)
    # iterate for iterate; at long horizons with tight curves a few per cent take one step more or less
    # (the two linear solvers round differently) and still end at the same point
    assert (g["iters"] == o["iters"])[ok].mean() >= (0.98 if N <= 31 else 0.9)
    assert np.abs(g["traj"] - o["traj"])[ok].max() <= 1e-5
    assert s.stats()["kernel_launches"] == 1
    wg, wo = g["traj"].copy(), g["traj"].copy()
    g2 = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"], warm=wg)
    o2 = oracle.solve_batch_frenet(ocfg, b["state"], b["kpoly"], b["v_des"], b["u_prev"], warm=wo, n_threads=8)
    _compare(g2, o2, min_conv=0.95)
    assert np.abs(wg - wo)[(g2["status"] == 0) & (o2["status"] == 0)].max() <= 1e-5
    assert g2["iters"].mean() < g["iters"].mean()


def test_frenet_full_size_properties(capi):
    """65,536 Frenet problems at N = 20: all converge; the returned trajectory satisfies the Frenet dynamics
    (:117-120), the bounds and the rate rows; solving again from the solution stays there."""
    N, B = 20, 65536
    s = capi.FrenetSolver(N)
    b = W.make_frenet_batch(B, N)
    g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    assert (g["status"] == 0).mean() >= 0.9999
    t = g["traj"]; n1 = N + 1
    sv, ey, v, ep, df, acc = t[:, :n1], t[:, n1:2 * n1], t[:, 2 * n1:3 * n1], t[:, 3 * n1:4 * n1], t[:, 4 * n1:4 * n1 + N], t[:, 4 * n1 + N:]
    k = b["kpoly"]
    K = ((k[:, 0:1] * sv[:, :N] + k[:, 1:2]) * sv[:, :N] + k[:, 2:3]) * sv[:, :N] + k[:, 3:4]
    bta = np.arctan(1.742 / (1.108 + 1.742) * np.tan(df))
    dsdt = v[:, :N] * np.cos(ep[:, :N] + bta) / (1.0 - ey[:, :N] * K)
    okm = g["status"] == 0
    assert np.abs(sv[:, 1:] - (sv[:, :N] + 0.2 * dsdt))[okm].max() <= 1e-7
    assert np.abs(ey[:, 1:] - (ey[:, :N] + 0.2 * v[:, :N] * np.sin(ep[:, :N] + bta)))[okm].max() <= 1e-7
    assert np.abs(ep[:, 1:] - (ep[:, :N] + 0.2 * (v[:, :N] / 1.742 * np.sin(bta) - dsdt * K)))[okm].max() <= 1e-7
    assert np.abs(v[:, 1:] - (v[:, :N] + 0.2 * acc))[okm].max() <= 1e-7
    assert np.abs(t[:, [0, n1, 3 * n1, 2 * n1]] - b["state"])[okm].max() <= 1e-7
    assert np.abs(acc).max() <= 1.0 and np.abs(df).max() <= 0.5 and v.min() >= 0.0 and v.max() <= 20.0
    assert np.abs(df[:, 0] - b["u_prev"][:, 0])[okm].max() <= 0.05 + 1e-7 and np.abs(acc[:, 0] - b["u_prev"][:, 1])[okm].max() <= 0.15 + 1e-7
    assert np.abs(np.diff(df[:, 1:], axis=1))[okm].max() <= 0.1 + 1e-7 and np.abs(np.diff(acc[:, 1:], axis=1))[okm].max() <= 0.3 + 1e-7
    w = t.copy()
    g2 = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"], warm=w)
    both = okm & (g2["status"] == 0)
    assert both.mean() >= 0.9999 and np.abs(g2["u0"] - g["u0"])[both].max() <= 1e-5


def test_frenet_module_mirror_and_errors(capi):
    """The API of MKZMPCPathFollowerFrenet.jl:132-207 as gazebo_sim_mpc_cmd_pub_frenet.jl:125-146 drives it."""
    from mkz_mpc_path_follower_b200.mpc_path_follower import MKZMPCPathFollowerFrenet
    kmpc = MKZMPCPathFollowerFrenet(N=8)
    b = W.make_frenet_batch(1, 8)
    kmpc.update_cost(9.0, 10.0, 0.5, 100.0, 1000.0, 0.0, 0.0)
    kmpc.update_init_cond(0.0, 0.0, float(b["state"][0, 2]), float(b["state"][0, 3]))
    kmpc.update_reference({"x": [0.0]}, b["kpoly"][0], float(b["v_des"][0]))
    a_opt, df_opt, is_opt = kmpc.solve_model()
    kmpc.update_current_input(df_opt, a_opt)
    res = kmpc.get_solver_results()
    assert is_opt == "Optimal" and len(res) == 8 and res[0].shape == (9,) and res[6].shape == (8,)
    assert res[7][0] == a_opt and res[6][0] == df_opt and abs(res[2][0] - b["state"][0, 3]) < 1e-7
    with pytest.raises(TypeError):
        kmpc.update_reference({}, np.zeros(3), 1.0)
    with pytest.raises(capi.MpcB200Error):
        capi.FrenetSolver(96)                     # horizons up to 95
    s = capi.FrenetSolver(8)
    with pytest.raises(capi.MpcB200Error):        # the XY entry point refuses a Frenet handle
        capi.Solver.solve_batch(s, np.zeros((1, 4)), np.zeros((1, 3, 9)), np.zeros((1, 2)))
    with pytest.raises(capi.MpcB200Error):        # ... and so does the closed-loop rollout
        s.rollout(np.zeros((1, 3)), np.zeros(1, dtype=np.int32), 2)


def test_line_search_failure_at_an_acceptable_point_gpu(capi, oracle):
    """The straggler of the 4-GPU rollout-start sweep (tests/test_emu_parity.py has the story): returns the
    point it had reached instead of restoring and running to the cap."""
    N = 20
    b = W.make_batch(1, N, b0=131072 + 127164)
    s = capi.Solver(N, start_mode=capi.START_ROLLOUT)
    g = s.solve_batch(b["state"], b["ref"], b["u_prev"], v_des=b["v_des"])
    o = oracle.solve_batch(_ocfg(oracle, s), b["state"], b["ref"], b["v_des"], b["u_prev"], n_threads=1)   # all-zero start
    assert g["status"][0] == 0 and g["iters"][0] <= 20
    assert np.abs(g["u0"] - o["u0"]).max() <= 1e-6 and abs(g["cost"][0] - o["cost"][0]) <= 1e-6 * o["cost"][0]


def test_frenet_closed_loop_holds_the_path(capi):
    """The Frenet module in closed loop (the control step of gazebo_sim_mpc_cmd_pub_frenet.jl:112-153 around the
    repository's plant): three vehicles start 0.7 m beside paths 1-3 with a heading error, speed up towards 8 m/s and
    settle on the path; every solve Optimal."""
    from mkz_mpc_path_follower_b200 import closed_loop
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    tabs = [GPSRefTrajectory(mat_filename=p).trajectory for p in (1, 2, 3)]
    pose0 = np.array([[tb[200, 4] - 0.7 * np.sin(tb[200, 3]), tb[200, 5] + 0.7 * np.cos(tb[200, 3]), tb[200, 3] + 0.08] for tb in tabs])   # 0.7 m to the left
    out = closed_loop.run_frenet([1, 2, 3], pose0, T=150, N=8)
    log = out["log"]
    assert (log[:, :, 6] == 0).mean() >= 0.995
    for b in range(3):
        err = closed_loop.path_errors(log[:, b:b + 1], tabs[b])[:, 0]
        assert err[0] > 0.4 and err[60:].max() < 0.25, (b, err[0], err[60:].max())
    assert log[-1, :, 3].min() > 4.0            # they do drive
    assert np.abs(log[:, :, 4]).max() <= 1.0 + 1e-9 and np.abs(log[:, :, 5]).max() <= 0.5 + 1e-9


def test_frenet_device_rollout_matches_host_loop(capi):
    """mpcb200_rollout_frenet (path ahead, curvature fit, warm-started Frenet solves and the plant in one persistent
    kernel) against closed_loop.run_frenet, which runs the same control step on the host around the same library."""
    from mkz_mpc_path_follower_b200 import closed_loop
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    N, T = 8, 40
    s = capi.FrenetSolver(N)
    trajs = [GPSRefTrajectory(mat_filename=p) for p in (1, 2, 3)]
    for i, g in enumerate(trajs):
        s.set_path(i, g.trajectory)
    rng = np.random.default_rng(12)
    path_ids, poses = [], []
    for p in (1, 2, 3):
        g = trajs[p - 1]
        for j in (200, 2600):
            nrm = np.array([-np.sin(g.trajectory[j, 3]), np.cos(g.trajectory[j, 3])])
            xy = g.trajectory[j, 4:6] + 0.7 * nrm
            poses.append((xy[0], xy[1], g.trajectory[j, 3] + rng.normal(scale=0.05))); path_ids.append(p)
    poses = np.array(poses); path_ids = np.array(path_ids)
    out = s.rollout(poses, (path_ids - 1).astype(np.int32), T, window=40.0, target_vel=8.0)
    assert s.stats()["kernel_launches"] == 1
    ref = closed_loop.run_frenet(path_ids, poses, T, N=N, window=40.0, target_vel=8.0)
    assert np.array_equal(out["log"][:, :, 6], ref["log"][:, :, 6]) and (out["log"][:, :, 6] == 0).all()
    assert np.abs(out["log"][:, :, 4:6] - ref["log"][:, :, 4:6]).max() <= 1e-5
    assert np.abs(out["log"][:, :, 0:4] - ref["log"][:, :, 0:4]).max() <= 1e-4
    for p in (1, 2, 3):
        err = closed_loop.path_errors(out["log"][-10:, path_ids == p, :], trajs[p - 1].trajectory)
        assert err.max() < 0.3          # started 0.7 m beside the path
    with pytest.raises(capi.MpcB200Error):
        capi.FrenetSolver(N).rollout(poses[:1], np.zeros(1, dtype=np.int32), 2)     # path not set
    with pytest.raises(capi.MpcB200Error):
        s.rollout(poses[:1], np.zeros(1, dtype=np.int32), 2, window=1.0)            # window too short for a cubic fit


@pytest.mark.parametrize("N,B", [(8, 48), (20, 32), (40, 8)])
def test_frenet_kkt_of_cuda_solutions(capi, oracle, N, B):
    """KKT conditions of the unscaled Frenet NLP at the CUDA solver's returned points (stress curvature), checked
    with the oracle's NLP functions only -- none of its interior-point code."""
    from test_frenet import frenet_kkt_check
    s = capi.FrenetSolver(N)
    b = _frenet_stress(W.make_frenet_batch(B, N), 23)
    g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"], want_traj=True)
    assert frenet_kkt_check(oracle, oracle.default_cfg_frenet(N), b, g["traj"], g["status"]) >= B - 2


def test_frenet_edge_cases(capi, oracle):
    """Empty batch, degenerate inputs (NaN curvature, singular 1 - e_y K, infeasible v0): the oracle's statuses."""
    from test_frenet import frenet_edge_batch
    N = 8
    s = capi.FrenetSolver(N, max_iter=60)
    g0 = s.solve_batch(np.zeros((0, 4)), np.zeros((0, 4)), np.zeros((0, 2)))
    assert g0["u0"].shape == (0, 2)
    b = frenet_edge_batch(N)
    g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"])
    o = oracle.solve_batch_frenet(oracle.default_cfg_frenet(N, max_iter=60), b["state"], b["kpoly"], b["v_des"], b["u_prev"], n_threads=1)
    assert (g["status"] == o["status"]).all() and g["status"].tolist() == [4, 4, 3, 1, 0, 0]
    ok = o["status"] == 0
    assert np.abs(g["u0"] - o["u0"])[ok].max() <= U_TOL
    with pytest.raises(ValueError):
        s.solve_batch(np.zeros((2, 4)), np.zeros((2, 3)), np.zeros((2, 2)))
    # the same degenerate problems (tiled past the small-batch path) through the thread-per-problem layout
    s.set_large_batch_path(1)
    rep = 12
    gt = s.solve_batch(np.tile(b["state"], (rep, 1)), np.tile(b["kpoly"], (rep, 1)), np.tile(b["u_prev"], (rep, 1)), v_des=np.tile(b["v_des"], rep))
    assert gt["status"].tolist() == [4, 4, 3, 1, 0, 0] * rep
    assert np.abs(gt["u0"][:6] - o["u0"])[ok].max() <= U_TOL and np.array_equal(gt["u0"][6:12][ok], gt["u0"][:6][ok])


@pytest.mark.parametrize("layout", ["warp", "thread"])
def test_frenet_large_batch_oracle_parity(capi, oracle, layout):
    """16,384 Frenet problems at N = 20, every one compared with the oracle: same status, the same iteration count
    on (nearly) all, |du| far inside the 1e-5 bar.  Both device layouts."""
    N, B = 20, 16384
    s = capi.FrenetSolver(N)
    s.set_large_batch_path(0 if layout == "warp" else 1)
    b = W.make_frenet_batch(B, N, b0=65536)
    g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"])
    o = oracle.solve_batch_frenet(oracle.default_cfg_frenet(N, tol=s.cfg.tol, max_iter=s.cfg.max_iter), b["state"], b["kpoly"],
                                  b["v_des"], b["u_prev"], n_threads=16)
    assert (g["status"] == o["status"]).all() and (g["status"] == 0).mean() >= 0.9999
    ok = o["status"] == 0
    assert (g["iters"] == o["iters"])[ok].mean() >= 0.999
    assert np.abs(g["u0"] - o["u0"])[ok].max() <= 1e-8
    assert (np.abs(g["cost"] - o["cost"])[ok] <= 1e-9 * np.maximum(1.0, o["cost"][ok])).all()


def test_handles_with_different_horizons_in_one_process(capi, oracle):
    """The dynamic shared-memory limit is an attribute of the kernel function, shared by every handle of the process: creating
    a short-horizon handle must not take the room a long-horizon handle needs (it did: `invalid argument` at the next launch of
    the N = 20 handle, which is what bench.py's strong-scaling key ran into on 2 GPUs)."""
    s20 = capi.Solver(20)
    b20 = W.make_batch(96, 20)
    g0 = s20.solve_batch(b20["state"], b20["ref"], b20["u_prev"], v_des=b20["v_des"])
    s8 = capi.Solver(8)                       # same kernels, smaller team footprint
    b8 = W.make_batch(96, 8)
    g8 = s8.solve_batch(b8["state"], b8["ref"], b8["u_prev"], v_des=b8["v_des"])
    s31 = capi.Solver(31)
    g1 = s20.solve_batch(b20["state"], b20["ref"], b20["u_prev"], v_des=b20["v_des"])
    assert (g0["status"] == g1["status"]).all() and np.array_equal(g0["u0"], g1["u0"]) and np.array_equal(g0["iters"], g1["iters"])
    o8 = oracle.solve_batch(_ocfg(oracle, s8), b8["state"], b8["ref"], b8["v_des"], b8["u_prev"], n_threads=4)
    _compare(g8, o8)
    b31 = W.make_batch(72, 31)
    s31.solve_batch(b31["state"], b31["ref"], b31["u_prev"], v_des=b31["v_des"])
    g2 = s8.solve_batch(b8["state"], b8["ref"], b8["u_prev"], v_des=b8["v_des"])
    assert np.array_equal(g8["u0"], g2["u0"])
