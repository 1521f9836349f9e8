"""Oracle NLP restatement: derivative checks (finite differences) and problem dimensions
(MKZMPCPathFollower.jl:65-123)."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _rand_point(O, cfg, rng):
    N = cfg.N
    n = 6 * N + 4
    z = rng.normal(size=n)
    for k in range(N + 1):
        z[6 * k + 3] = rng.uniform(0.5, 15.0)  # v
        z[6 * k + 2] = rng.uniform(-3.0, 3.0)  # psi
        if k < N:
            z[6 * k + 4] = rng.uniform(-1, 1)
            z[6 * k + 5] = rng.uniform(-0.5, 0.5)
    return z


@pytest.mark.parametrize("N", [8, 20])
def test_dimensions(oracle, N):
    cfg = oracle.default_cfg(N)
    L = oracle.lib()
    # N=8: 52 vars, 36 eq, 14 range rows; N=20: 124 / 84 / 38
    assert L.mpc_oracle_nvar(C.byref(cfg)) == 6 * N + 4
    assert L.mpc_oracle_ncon(C.byref(cfg)) == 4 * N + 4
    assert L.mpc_oracle_nrange(C.byref(cfg)) == 2 * (N - 1)
    assert {8: (52, 36, 14), 20: (124, 84, 38)}[N] == (6 * N + 4, 4 * N + 4, 2 * (N - 1))


@pytest.mark.parametrize("N", [4, 8])
def test_derivatives_fd(oracle, N):
    rng = np.random.default_rng(3)
    cfg = oracle.default_cfg(N, weights=[9, 9, 10, 2.0, 100, 1000, 0.3, 0.7])
    L = oracle.lib()
    n, mc, md = 6 * N + 4, 4 * N + 4, 2 * (N - 1)
    ref = rng.normal(size=3 * (N + 1))
    state = rng.normal(size=4)
    u_prev = np.array([0.01, 0.2])
    z = _rand_point(oracle, cfg, rng)
    g = np.empty(n)
    L.mpc_oracle_eval_grad_f(C.byref(cfg), _p(ref), 1.3, _p(z), _p(g))
    Jc = np.empty((mc, n)); Jd = np.empty((md, n))
    L.mpc_oracle_eval_jac(C.byref(cfg), _p(z), _p(Jc), _p(Jd))
    yc = rng.normal(size=mc)
    H = np.empty((n, n))
    L.mpc_oracle_eval_hess(C.byref(cfg), _p(z), 0.7, _p(yc), _p(H))
    assert np.allclose(H, H.T)

    def f(zz):
        return L.mpc_oracle_eval_f(C.byref(cfg), _p(ref), 1.3, _p(zz))

    def c(zz):
        o = np.empty(mc); L.mpc_oracle_eval_c(C.byref(cfg), _p(state), _p(zz), _p(o)); return o

    def d(zz):
        o = np.empty(md); L.mpc_oracle_eval_d(C.byref(cfg), _p(u_prev), _p(zz), _p(o)); return o

    def lag_grad(zz):
        gg = np.empty(n); L.mpc_oracle_eval_grad_f(C.byref(cfg), _p(ref), 1.3, _p(zz), _p(gg))
        J1 = np.empty((mc, n)); J2 = np.empty((md, n))
        L.mpc_oracle_eval_jac(C.byref(cfg), _p(zz), _p(J1), _p(J2))
        return 0.7 * gg + J1.T @ yc

    h = 1e-6
    g_fd = np.empty(n); Jc_fd = np.empty((mc, n)); Jd_fd = np.empty((md, n)); H_fd = np.empty((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = h
        g_fd[i] = (f(z + e) - f(z - e)) / (2 * h)
        Jc_fd[:, i] = (c(z + e) - c(z - e)) / (2 * h)
        Jd_fd[:, i] = (d(z + e) - d(z - e)) / (2 * h)
        H_fd[:, i] = (lag_grad(z + e) - lag_grad(z - e)) / (2 * h)
    assert np.allclose(g, g_fd, rtol=1e-6, atol=1e-5)
    assert np.allclose(Jc, Jc_fd, rtol=1e-6, atol=1e-7)
    assert np.allclose(Jd, Jd_fd, rtol=1e-6, atol=1e-7)
    assert np.allclose(H, H_fd, rtol=1e-5, atol=1e-5)


def test_quirks(oracle):
    """Q1: pair (1,0) has no rate row; Q2: terminal speed uncosted; Q6: ref index 0 unused."""
    N = 8
    cfg = oracle.default_cfg(N, weights=[9, 9, 10, 5.0, 100, 1000, 0, 0])
    L = oracle.lib()
    n, md = 6 * N + 4, 2 * (N - 1)
    Jc = np.empty((4 * N + 4, n)); Jd = np.empty((md, n))
    z = np.zeros(n)
    L.mpc_oracle_eval_jac(C.byref(cfg), _p(z), _p(Jc), _p(Jd))
    # first-move rows touch only stage 0; no row couples stage 0 with stage 1
    rows_touching_u0 = [r for r in range(md) if Jd[r, 4] != 0 or Jd[r, 5] != 0]
    assert rows_touching_u0 == [0, 1]
    rows_touching_u1 = [r for r in range(md) if Jd[r, 6 + 4] != 0 or Jd[r, 6 + 5] != 0]
    assert rows_touching_u1 == [2, 3]  # pair (2,1)
    ref = np.zeros(3 * (N + 1))
    f0 = L.mpc_oracle_eval_f(C.byref(cfg), _p(ref), 0.0, _p(z))
    z2 = z.copy(); z2[6 * N + 3] = 7.0  # terminal v
    assert L.mpc_oracle_eval_f(C.byref(cfg), _p(ref), 0.0, _p(z2)) == f0
    z3 = z.copy(); z3[3] = 7.0  # v_0
    assert L.mpc_oracle_eval_f(C.byref(cfg), _p(ref), 0.0, _p(z3)) == f0
    ref2 = ref.copy(); ref2[0] = 5.0; ref2[N + 1] = 5.0; ref2[2 * (N + 1)] = 5.0
    assert L.mpc_oracle_eval_f(C.byref(cfg), _p(ref2), 0.0, _p(z)) == f0
