"""Host-side mirror of the reference's Julia module `MKZMPCPathFollower`
(scripts/mpc_utils/MKZMPCPathFollower.jl:132-207): same function names, argument order and
return tuples, backed by libmpc_b200.so with a batch of one.

    kmpc = MKZMPCPathFollower()                       # import MKZMPCPathFollower (:23-128)
    kmpc.update_cost(9., 9., 10., 0., 100., 1000., 0., 0.)
    kmpc.update_init_cond(x, y, psi, v)
    kmpc.update_reference(x_ref, y_ref, psi_ref, des_speed)
    a_opt, df_opt, is_opt = kmpc.solve_model()        # acceleration first
    kmpc.update_current_input(df_opt, a_opt)          # steering first
    res = kmpc.get_solver_results()                   # x, y, v, psi, x_ref, y_ref, psi_ref, d_f, acc

Like the reference (JuMP <= 0.18 re-solves from the previous primal solution) each solve_model()
starts from the previous solution; the very first one starts from the module-load solve of the
default problem (:36-39, :126-128).  Failure is reported through the returned status symbol,
never by raising (mpc_cmd_pub.jl:121-132 publishes whatever comes back).
"""
import numpy as np

from . import capi


class MKZMPCPathFollower(object):
    def __init__(self, N=8, device=0, **cfg_overrides):
        self._solver = capi.Solver(N=N, device=device, **cfg_overrides)
        cfg = self._solver.cfg
        self.N = cfg.N          # module constants readable as kmpc.N, kmpc.dt (mpc_cmd_pub.jl:51)
        self.dt = cfg.dt
        self.dt_control = cfg.dt_control
        self.L_a, self.L_b = cfg.L_a, cfg.L_b
        N = self.N
        # parameter defaults of MKZMPCPathFollower.jl:36-39,75,82,110-113
        v_ref = 15.0
        self._x_ref = v_ref * np.arange(N + 1) * self.dt
        self._y_ref = np.zeros(N + 1)
        self._psi_ref = np.zeros(N + 1)
        self._v_target = v_ref
        self._state = np.zeros(4)
        self._u_curr = np.zeros(2)  # d_f_current, acc_current
        self._warm = np.zeros((1, 6 * N + 4))  # start=0.0 (:65-72)
        self._last = None
        self.solve_model()  # "MPC: Initial solve ..." (:126-128)

    # ---- MKZMPCPathFollower.jl:132-138 ----
    def update_init_cond(self, x, y, psi, vel):
        self._state[:] = (float(x), float(y), float(psi), float(vel))

    # ---- :142-147 ----
    def update_reference(self, x_ref, y_ref, psi_ref, v_des):
        N = self.N
        x_ref = np.asarray(x_ref, dtype=np.float64); y_ref = np.asarray(y_ref, dtype=np.float64)
        psi_ref = np.asarray(psi_ref, dtype=np.float64)
        if x_ref.shape != (N + 1,) or y_ref.shape != (N + 1,) or psi_ref.shape != (N + 1,):
            raise TypeError("update_reference: expected three Float64 arrays of length N+1")  # Julia MethodError
        self._x_ref, self._y_ref, self._psi_ref = x_ref.copy(), y_ref.copy(), psi_ref.copy()
        self._v_target = float(v_des)

    # ---- :151-154  (steering first!) ----
    def update_current_input(self, c_swa, c_acc):
        self._u_curr[:] = (float(c_swa), float(c_acc))

    # ---- :158-169 ----
    def update_cost(self, cx, cy, cp, cv, cda, cdd, ca, cd):
        self._solver.set_cost([cx, cy, cp, cv, cda, cdd, ca, cd])

    # ---- :173-183 ----
    def solve_model(self):
        ref = np.stack((self._x_ref, self._y_ref, self._psi_ref))[None]
        out = self._solver.solve_batch(self._state[None], ref, self._u_curr[None], v_des=np.array([self._v_target]),
                                       warm=self._warm, want_traj=False)
        self._last = out
        acc_opt, d_f_opt = float(out["u0"][0, 0]), float(out["u0"][0, 1])
        return acc_opt, d_f_opt, capi.STATUS_SYMBOLS[int(out["status"][0])]

    # ---- :188-207 ----
    def get_solver_results(self):
        N = self.N
        t = self._warm[0]
        x_mpc, y_mpc, v_mpc, psi_mpc = (t[0:N + 1].copy(), t[N + 1:2 * (N + 1)].copy(),
                                        t[2 * (N + 1):3 * (N + 1)].copy(), t[3 * (N + 1):4 * (N + 1)].copy())
        d_f_opt = t[4 * (N + 1):4 * (N + 1) + N].copy()
        acc_opt = t[4 * (N + 1) + N:].copy()
        return (x_mpc, y_mpc, v_mpc, psi_mpc, self._x_ref.copy(), self._y_ref.copy(), self._psi_ref.copy(),
                d_f_opt, acc_opt)

    @property
    def last_iters(self):
        return None if self._last is None else int(self._last["iters"][0])

    @property
    def last_cost(self):
        return None if self._last is None else float(self._last["cost"][0])


class MKZMPCPathFollowerFrenet(object):
    """Mirror of the reference's Frenet-frame module `MKZMPCPathFollowerFrenet`
    (scripts/mpc_utils/MKZMPCPathFollowerFrenet.jl:132-207), as the Gazebo node drives it
    (scripts/nodes_gazebo_sim/gazebo_sim_mpc_cmd_pub_frenet.jl:112-153):

        kmpc.update_init_cond(0.0, 0.0, -des_init_heading, curr_speed)   # s, ey, epsi, v
        kmpc.update_reference(path_ref, K_coeffs, des_speed)
        a_opt, df_opt, is_opt = kmpc.solve_model()
        kmpc.update_current_input(df_opt, a_opt)
        res = kmpc.get_solver_results()    # s, ey, v, epsi, K, path_ref, d_f, acc
    """

    def __init__(self, N=8, device=0, **cfg_overrides):
        self._solver = capi.FrenetSolver(N=N, device=device, **cfg_overrides)
        cfg = self._solver.cfg
        self.N, self.dt, self.dt_control = cfg.N, cfg.dt, cfg.dt_control
        self.L_a, self.L_b = cfg.L_a, cfg.L_b
        self._k = np.zeros(4)        # k_coeff_ref (:40-41)
        self._v_target = 15.0        # v_ref (:39)
        self._path_ref = {}          # path_ref (:38)
        self._state = np.zeros(4)
        self._u_curr = np.zeros(2)
        self._warm = np.zeros((1, 6 * self.N + 4))   # start=0.0 (:65-73)
        self._last = None
        self.solve_model()           # "MPC: Initial solve ..." (:126-128)

    def update_init_cond(self, s, ey, epsi, vel):          # :132-138
        self._state[:] = (float(s), float(ey), float(epsi), float(vel))

    def update_reference(self, path, k_coeffs, v_des):     # :142-147
        k = np.asarray(k_coeffs, dtype=np.float64)
        if k.shape != (4,):
            raise TypeError("update_reference: k_coeffs must be a Float64 array of length 4")
        self._path_ref = path
        self._k = k.copy()
        self._v_target = float(v_des)

    def update_current_input(self, c_swa, c_acc):          # :151-154 (steering first)
        self._u_curr[:] = (float(c_swa), float(c_acc))

    def update_cost(self, cey, cep, cev, cda, cdd, ca, cd):   # :158-169
        self._solver.set_cost([cey, cep, cev, cda, cdd, ca, cd])

    def solve_model(self):                                 # :173-183
        out = self._solver.solve_batch(self._state[None], self._k[None], self._u_curr[None], v_des=np.array([self._v_target]),
                                       warm=self._warm)
        self._last = out
        return float(out["u0"][0, 0]), float(out["u0"][0, 1]), capi.STATUS_SYMBOLS[int(out["status"][0])]

    def get_solver_results(self):                          # :188-207
        N = self.N
        t = self._warm[0]
        s_mpc, ey_mpc, v_mpc, epsi_mpc = (t[0:N + 1].copy(), t[N + 1:2 * (N + 1)].copy(),
                                          t[2 * (N + 1):3 * (N + 1)].copy(), t[3 * (N + 1):4 * (N + 1)].copy())
        return (s_mpc, ey_mpc, v_mpc, epsi_mpc, self._k.copy(), self._path_ref,
                t[4 * (N + 1):4 * (N + 1) + N].copy(), t[4 * (N + 1) + N:].copy())
