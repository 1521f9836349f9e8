"""Oracle interior-point solver, validated intrinsically (the reference holds no golden
solver outputs -- parity unpinned):
  * the returned point satisfies the KKT conditions of the UNSCALED reference NLP,
  * an independent solver (scipy SLSQP with the analytic derivatives) reaches the same point.
"""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import minimize

from mkz_mpc_path_follower_b200 import workload as W


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _nlp(O, cfg, state, ref, v_des, u_prev):
    L = O.lib()
    N = cfg.N
    n, mc, md = 6 * N + 4, 4 * N + 4, 2 * (N - 1)
    ref = np.ascontiguousarray(ref).reshape(-1)

    def f(z):
        return L.mpc_oracle_eval_f(C.byref(cfg), _p(ref), v_des, _p(np.ascontiguousarray(z)))

    def g(z):
        o = np.empty(n); L.mpc_oracle_eval_grad_f(C.byref(cfg), _p(ref), v_des, _p(np.ascontiguousarray(z)), _p(o)); return o

    def c(z):
        o = np.empty(mc); L.mpc_oracle_eval_c(C.byref(cfg), _p(state), _p(np.ascontiguousarray(z)), _p(o)); return o

    def d(z):
        o = np.empty(md); L.mpc_oracle_eval_d(C.byref(cfg), _p(u_prev), _p(np.ascontiguousarray(z)), _p(o)); return o

    def jac(z):
        Jc = np.empty((mc, n)); Jd = np.empty((md, n))
        L.mpc_oracle_eval_jac(C.byref(cfg), _p(np.ascontiguousarray(z)), _p(Jc), _p(Jd)); return Jc, Jd

    lo = np.full(n, -np.inf); hi = np.full(n, np.inf)
    for k in range(N + 1):
        lo[6 * k + 3], hi[6 * k + 3] = cfg.v_min, cfg.v_max
        if k < N:
            lo[6 * k + 4], hi[6 * k + 4] = -cfg.a_max, cfg.a_max
            lo[6 * k + 5], hi[6 * k + 5] = -cfg.steer_max, cfg.steer_max
    dlim = np.empty(md)
    for r in range(md):
        h = cfg.dt_control if r // 2 == 0 else cfg.dt
        dlim[r] = (cfg.steer_dmax if r % 2 == 0 else cfg.a_dmax) * h
    return f, g, c, d, jac, lo, hi, dlim


def _kkt_residual(O, cfg, prob, z):
    """Least-squares multipliers on the active set; returns (stationarity residual, primal
    violation, worst multiplier sign violation)."""
    f, g, c, d, jac, lo, hi, dlim = _nlp(O, cfg, *prob)
    Jc, Jd = jac(z)
    n = z.size
    # interior-point solutions approach weakly active bounds like sqrt(mu): use a generous
    # activity threshold and let the sign check reject wrong multipliers
    tol_act = 1e-3
    cols = [Jc.T]
    signs = []
    for i in range(n):
        if np.isfinite(lo[i]) and z[i] - lo[i] < tol_act:
            e = np.zeros((n, 1)); e[i] = -1.0; cols.append(e); signs.append(len(signs))
        if np.isfinite(hi[i]) and hi[i] - z[i] < tol_act:
            e = np.zeros((n, 1)); e[i] = 1.0; cols.append(e); signs.append(len(signs))
    dv = d(z)
    for r in range(dv.size):
        if dv[r] + dlim[r] < tol_act:
            cols.append(-Jd[r][:, None]); signs.append(len(signs))
        if dlim[r] - dv[r] < tol_act:
            cols.append(Jd[r][:, None]); signs.append(len(signs))
    A = np.hstack(cols)
    sol, *_ = np.linalg.lstsq(A, -g(z), rcond=None)
    stat = np.abs(A @ sol + g(z)).max()
    lam_ineq = sol[Jc.shape[0]:]
    sign_viol = max(0.0, -lam_ineq.min()) if lam_ineq.size else 0.0
    prim = max(np.abs(c(z)).max(), np.maximum(lo - z, 0).max(), np.maximum(z - hi, 0).max(),
               np.maximum(np.abs(dv) - dlim, 0).max())
    return stat, prim, sign_viol


@pytest.mark.parametrize("N,B", [(8, 24), (20, 8)])
def test_kkt_of_returned_point(oracle, N, B):
    cfg = oracle.default_cfg(N)
    b = W.make_batch(B, N)
    nconv = 0
    for j in range(B):
        r = oracle.solve(cfg, b["state"][j], b["ref"][j], 1.0, b["u_prev"][j])
        if r["status"] != 0:
            continue
        nconv += 1
        z = np.empty(6 * N + 4)
        oracle.lib().mpc_oracle_traj_to_z(C.byref(cfg), _p(r["traj"]), _p(z))
        stat, prim, sign = _kkt_residual(oracle, cfg, (b["state"][j], b["ref"][j], 1.0, b["u_prev"][j]), z)
        gscale = max(1.0, np.abs(z).max())
        assert prim <= 1e-8 * gscale, (j, prim)
        assert stat <= 1e-4, (j, stat)       # unscaled; Ipopt tol 1e-8 acts on the scaled problem
        assert sign <= 1e-6, (j, sign)
    assert nconv >= int(0.85 * B)


@pytest.mark.parametrize("N", [8])
def test_agrees_with_scipy_slsqp(oracle, N):
    """Independent solver, analytic derivatives, started from the oracle's own warm point
    perturbed -- checks that the oracle's point is the local minimiser SLSQP also finds."""
    cfg = oracle.default_cfg(N)
    b = W.make_batch(6, N)
    rng = np.random.default_rng(0)
    checked = 0
    for j in range(6):
        prob = (b["state"][j], b["ref"][j], 1.0, b["u_prev"][j])
        r = oracle.solve(cfg, *prob)
        if r["status"] != 0:
            continue
        f, g, c, d, jac, lo, hi, dlim = _nlp(oracle, cfg, *prob)
        z0 = np.empty(6 * N + 4)
        oracle.lib().mpc_oracle_traj_to_z(C.byref(cfg), _p(r["traj"]), _p(z0))
        zs = z0 + 1e-2 * rng.normal(size=z0.size)
        cons = [
            {"type": "eq", "fun": c, "jac": lambda z: jac(z)[0]},
            {"type": "ineq", "fun": lambda z: dlim - d(z), "jac": lambda z: -jac(z)[1]},
            {"type": "ineq", "fun": lambda z: dlim + d(z), "jac": lambda z: jac(z)[1]},
        ]
        res = minimize(f, zs, jac=g, bounds=list(zip(lo, hi)), constraints=cons, method="SLSQP",
                       options={"ftol": 1e-15, "maxiter": 500})
        assert np.abs(c(res.x)).max() < 1e-7
        # same minimiser: cost to 1e-6 relative, first move to 1e-4 (SLSQP's own accuracy)
        assert abs(res.fun - r["cost"]) <= 1e-6 * max(1.0, abs(r["cost"])), (j, res.fun, r["cost"])
        assert np.abs(res.x[4:6] - z0[4:6]).max() < 1e-4, (j, res.x[4:6], z0[4:6])
        checked += 1
    assert checked >= 4


def test_warm_start_same_optimum(oracle):
    N = 8
    cfg = oracle.default_cfg(N)
    b = W.make_batch(32, N)
    cold = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=4)
    warm = cold["traj"].copy()
    again = oracle.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=warm, n_threads=4)
    ok = (cold["status"] == 0) & (again["status"] == 0)
    assert ok.sum() >= 28
    # two different iterate paths to the same KKT point: tolerance of the stopping rule, not rounding
    assert np.abs(cold["u0"] - again["u0"])[ok].max() < 1e-5
    assert (np.abs(cold["cost"] - again["cost"])[ok] <= 1e-6 * np.maximum(1.0, np.abs(cold["cost"][ok]))).all()
    assert again["iters"][ok].mean() < 15


def test_infeasible_inputs(oracle):
    N = 8
    cfg = oracle.default_cfg(N)
    b = W.make_batch(1, N)
    st = b["state"][0].copy(); st[3] = -0.5
    r = oracle.solve(cfg, st, b["ref"][0], 1.0, b["u_prev"][0])
    assert r["status"] == 1
    r = oracle.solve(cfg, b["state"][0], b["ref"][0], 1.0, np.array([0.7, 0.0]))
    assert r["status"] == 1


def test_iteration_cap_is_userlimit(oracle):
    N = 8
    cfg = oracle.default_cfg(N, max_iter=2)
    b = W.make_batch(1, N)
    r = oracle.solve(cfg, b["state"][0], b["ref"][0], 1.0, b["u_prev"][0])
    assert r["status"] == 3 and r["iters"] == 2
