#!/usr/bin/env python
"""Golden solver outputs from a REAL Ipopt, if one is reachable.

The reference solves its NLP with JuMP <= 0.18 + Ipopt.jl (MKZMPCPathFollower.jl:29,176); neither Julia nor Ipopt exists
in the build container or, so far, on the GPU box.  bench.py probes for them at run time (`ipopt_probe`); if `cyipopt`
is importable this module solves the first problems of the bench batch with it -- the NLP of MKZMPCPathFollower.jl:65-123
through the oracle's evaluation callbacks (objective, gradient, equality rows c, range rows d, Jacobian, Lagrangian
Hessian), Ipopt options as the reference leaves them (defaults, print_level = 0; max_cpu_time is NOT set so that the
vectors are reproducible) -- and dumps states, references, statuses, first moves and costs as an .npz that
tests/test_ipopt_golden.py picks up when present.  TEST INFRASTRUCTURE: nothing in the product imports this.

    python tools/ipopt_golden.py --selftest        # exercises the callbacks without cyipopt
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class MpcNlp(object):
    """cyipopt problem object for one problem of the batch.  Variables in the oracle's internal order
    (x, y, psi, v, acc, df per stage, then the terminal state), constraints = [c (4 + 4N equalities); d (2(N-1) range rows)]."""

    def __init__(self, O, cfg, state, ref, v_des, u_prev):
        self.O, self.cfg, self.L = O, cfg, O.lib()
        self.state = np.ascontiguousarray(state, dtype=np.float64)
        self.ref = np.ascontiguousarray(ref, dtype=np.float64).reshape(-1)
        self.u_prev = np.ascontiguousarray(u_prev, dtype=np.float64)
        self.v_des = float(v_des)
        self.n = self.L.mpc_oracle_nvar(C.byref(cfg)); self.mc = self.L.mpc_oracle_ncon(C.byref(cfg)); self.md = self.L.mpc_oracle_nrange(C.byref(cfg))
        self.hrow, self.hcol = np.tril_indices(self.n)

    def _p(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def bounds(self):
        cfg, N = self.cfg, self.cfg.N
        lb = np.full(self.n, -2e19); ub = np.full(self.n, 2e19)
        for k in range(N + 1):
            lb[6 * k + 3], ub[6 * k + 3] = cfg.v_min, cfg.v_max
            if k < N:
                lb[6 * k + 4], ub[6 * k + 4] = -cfg.a_max, cfg.a_max
                lb[6 * k + 5], ub[6 * k + 5] = -cfg.steer_max, cfg.steer_max
        cl = np.zeros(self.mc + self.md); cu = np.zeros(self.mc + self.md)
        # range rows in the oracle's order (see mpc_oracle_eval_d): first move vs previous command with dt_control,
        # then pairs (q+1, q), q = 1..N-2, with dt; steering row then acceleration row
        lim = []
        for q in range(N - 1):
            h = cfg.dt_control if q == 0 else cfg.dt
            lim += [cfg.steer_dmax * h, cfg.a_dmax * h]
        cl[self.mc:] = -np.array(lim); cu[self.mc:] = np.array(lim)
        return lb, ub, cl, cu

    def objective(self, z):
        z = np.ascontiguousarray(z)
        return self.L.mpc_oracle_eval_f(C.byref(self.cfg), self._p(self.ref), self.v_des, self._p(z))

    def gradient(self, z):
        z = np.ascontiguousarray(z); g = np.empty(self.n)
        self.L.mpc_oracle_eval_grad_f(C.byref(self.cfg), self._p(self.ref), self.v_des, self._p(z), self._p(g))
        return g

    def constraints(self, z):
        z = np.ascontiguousarray(z); c = np.empty(self.mc); d = np.empty(self.md)
        self.L.mpc_oracle_eval_c(C.byref(self.cfg), self._p(self.state), self._p(z), self._p(c))
        self.L.mpc_oracle_eval_d(C.byref(self.cfg), self._p(self.u_prev), self._p(z), self._p(d))
        return np.concatenate((c, d))

    def jacobian(self, z):   # dense, row-major
        z = np.ascontiguousarray(z); Jc = np.empty((self.mc, self.n)); Jd = np.empty((self.md, self.n))
        self.L.mpc_oracle_eval_jac(C.byref(self.cfg), self._p(z), self._p(Jc), self._p(Jd))
        return np.concatenate((Jc, Jd)).reshape(-1)

    def hessianstructure(self):
        return self.hrow, self.hcol

    def hessian(self, z, lagrange, obj_factor):
        z = np.ascontiguousarray(z); yc = np.ascontiguousarray(lagrange[:self.mc], dtype=np.float64); H = np.empty((self.n, self.n))
        self.L.mpc_oracle_eval_hess(C.byref(self.cfg), self._p(z), float(obj_factor), self._p(yc), self._p(H))
        return H[self.hrow, self.hcol]


def _oracle():
    from oracle import oracle as O
    O.build()
    L = O.lib()
    L.mpc_oracle_eval_f.restype = C.c_double
    return O


def dump(path, N=20, n=256):
    import cyipopt
    from mkz_mpc_path_follower_b200 import workload
    O = _oracle()
    cfg = O.default_cfg(N)
    b = workload.make_batch(n, N)
    u0 = np.zeros((n, 2)); cost = np.zeros(n); status = np.zeros(n, dtype=np.int32); iters = np.zeros(n, dtype=np.int32)
    traj = np.zeros((n, 6 * N + 4))
    for i in range(n):
        nlp = MpcNlp(O, cfg, b["state"][i], b["ref"][i], b["v_des"][i], b["u_prev"][i])
        lb, ub, cl, cu = nlp.bounds()
        prob = cyipopt.Problem(n=nlp.n, m=nlp.mc + nlp.md, problem_obj=nlp, lb=lb, ub=ub, cl=cl, cu=cu)
        prob.add_option("print_level", 0)
        z, info = prob.solve(np.zeros(nlp.n))          # start = 0.0 (MKZMPCPathFollower.jl:65-72)
        status[i] = info["status"]; cost[i] = info["obj_val"]
        u0[i] = (z[4], z[5])                            # acc_opt[1], d_f_opt[1]
        t = np.empty(6 * N + 4)
        O.lib().mpc_oracle_z_to_traj(C.byref(cfg), nlp._p(np.ascontiguousarray(z)), nlp._p(t))
        traj[i] = t
    np.savez_compressed(path, N=N, state=b["state"], ref=b["ref"], v_des=b["v_des"], u_prev=b["u_prev"],
                        ipopt_status=status, u0=u0, cost=cost, traj=traj, cyipopt_version=str(getattr(cyipopt, "__version__", "?")))
    return {"file": path, "problems": n, "ipopt_success": int((status == 0).sum()), "ipopt_acceptable": int((status == 1).sum())}


def selftest():
    from mkz_mpc_path_follower_b200 import workload
    O = _oracle()
    N = 8
    cfg = O.default_cfg(N)
    b = workload.make_batch(2, N)
    nlp = MpcNlp(O, cfg, b["state"][0], b["ref"][0], b["v_des"][0], b["u_prev"][0])
    lb, ub, cl, cu = nlp.bounds()
    r = O.solve(cfg, b["state"][0], b["ref"][0], b["v_des"][0], b["u_prev"][0])
    z = np.empty(nlp.n)
    O.lib().mpc_oracle_traj_to_z(C.byref(cfg), nlp._p(np.ascontiguousarray(r["traj"])), nlp._p(z))
    g = nlp.constraints(z)
    assert abs(nlp.objective(z) - r["cost"]) <= 1e-9 * max(1.0, abs(r["cost"]))
    assert np.abs(g[:nlp.mc]).max() <= 1e-7 and (g[nlp.mc:] >= cl[nlp.mc:] - 1e-7).all() and (g[nlp.mc:] <= cu[nlp.mc:] + 1e-7).all()
    assert (z >= lb - 1e-7).all() and (z <= ub + 1e-7).all()
    assert nlp.gradient(z).shape == (nlp.n,) and nlp.jacobian(z).shape == ((nlp.mc + nlp.md) * nlp.n,)
    assert nlp.hessian(z, np.ones(nlp.mc + nlp.md), 1.0).shape == nlp.hrow.shape
    print("ipopt_golden selftest ok: n = %d, m = %d + %d; the oracle's optimum is feasible for the cyipopt transcription" % (nlp.n, nlp.mc, nlp.md))


if __name__ == "__main__":
    if "--selftest" in sys.argv:
        selftest()
    else:
        print(dump(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ipopt_N20.npz")))
