"""Batched restatement of the reference plant, scripts/vehicle_simulator.py:58-112: dynamic bicycle
with a linear tyre model, 10 explicit-Euler sub-steps of 1 ms per 100 Hz publish, first-order
actuator lag.  State per vehicle: X, Y, psi, vx, vy, wz, acc, df (the fields the node publishes as
`state_est` are x=X, y=Y, psi, v=vx, a=acc, df; vehicle_simulator.py:41-50).

Host-side (numpy) harness code: the closed-loop Monte-Carlo rollout itself runs on the GPU through
mpcb200_rollout; this class is what the Python harness and the tests use to step single vehicles or
small batches exactly like the ROS node does.
"""
import numpy as np

# vehicle_simulator.py:61-67
LF, LR, M, IZ, C_ALPHA_F, C_ALPHA_R = 1.152, 1.693, 1840, 3477, 4.0703e4, 6.4495e4
# The reference node is Python 2 and `m = 1840` is an int, so `1/m` in the vx equation (:86) is integer
# division = 0: the tyre-force term drops out of the longitudinal equation AS THE REFERENCE RUNS (the vy and
# wz equations, :90-91, use 1.0/m and 1.0/Iz and keep theirs).  Restated as it runs.
INV_M_VX = 1 // M


class VehicleSimulator(object):
    def __init__(self, X0=-300.0, Y0=-450.0, Psi0=1.0, batch=None):
        """Defaults are the node's rosparam defaults (vehicle_simulator.py:28-30).  With `batch`
        the initial pose arguments may be arrays of that length."""
        n = 1 if batch is None else int(batch)
        self._scalar = batch is None
        self.X = np.broadcast_to(np.asarray(X0, dtype=np.float64), (n,)).copy()
        self.Y = np.broadcast_to(np.asarray(Y0, dtype=np.float64), (n,)).copy()
        self.psi = np.broadcast_to(np.asarray(Psi0, dtype=np.float64), (n,)).copy()
        self.vx = np.zeros(n); self.vy = np.zeros(n); self.wz = np.zeros(n)
        self.acc = np.zeros(n); self.df = np.zeros(n)
        self.acc_des = np.zeros(n); self.df_des = np.zeros(n)
        self.dt_model = 0.01

    def mpc_cmd(self, accel_cmd, steer_angle_cmd):   # _mpc_cmd_callback, :53-56
        self.acc_des[:] = accel_cmd
        self.df_des[:] = steer_angle_cmd

    def update_vehicle_model(self, disc_steps=10):   # _update_vehicle_model, :58-106
        deltaT = self.dt_model / disc_steps
        for _ in range(disc_steps):
            moving = np.fabs(self.vx) > 1e-6
            vxs = np.where(moving, self.vx, 1.0)
            alpha_f = np.where(moving, self.df - np.arctan2(self.vy + LF * self.wz, vxs), 0.0)
            alpha_r = np.where(moving, -np.arctan2(self.vy - LF * self.wz, vxs), 0.0)   # :77 uses lf (sic)
            Fyf = C_ALPHA_F * alpha_f
            Fyr = C_ALPHA_R * alpha_r
            vx_n = np.maximum(0.0, self.vx + deltaT * (self.acc - INV_M_VX * Fyf * np.sin(self.df) + self.wz * self.vy))
            fwd = vx_n > 1e-6
            vy_n = np.where(fwd, self.vy + deltaT * (1.0 / M * (Fyf * np.cos(self.df) + Fyr) - self.wz * self.vx), 0.0)
            wz_n = np.where(fwd, self.wz + deltaT * (1.0 / IZ * (LF * Fyf * np.cos(self.df) - LR * Fyr)), 0.0)
            psi_n = self.psi + deltaT * self.wz
            X_n = self.X + deltaT * (self.vx * np.cos(self.psi) - self.vy * np.sin(self.psi))
            Y_n = self.Y + deltaT * (self.vx * np.sin(self.psi) + self.vy * np.cos(self.psi))
            self.X, self.Y = X_n, Y_n
            self.psi = (psi_n + np.pi) % (2.0 * np.pi) - np.pi
            self.vx, self.vy, self.wz = vx_n, vy_n, wz_n
            # _update_low_level_control, :108-112
            self.acc = 5.0 * (self.acc_des - self.acc) * deltaT + self.acc
            self.df = 5.0 * (self.df_des - self.df) * deltaT + self.df

    def state_est(self):
        """x, y, psi, v, a, df as published (vehicle_simulator.py:41-50); (B,6), or (6,) unbatched."""
        s = np.stack((self.X, self.Y, self.psi, self.vx, self.acc, self.df), axis=1)
        return s[0] if self._scalar else s

    def full_state(self):
        s = np.stack((self.X, self.Y, self.psi, self.vx, self.vy, self.wz, self.acc, self.df), axis=1)
        return s[0] if self._scalar else s
