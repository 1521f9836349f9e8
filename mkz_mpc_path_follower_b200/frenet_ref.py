"""Frenet-frame reference: curvature polynomial K(s) fitted to a waypoint window.

Restates scripts/sim_path_utils/nav_msgs_path_frenet.py (reference file:line below) without
rospy / matplotlib: cubic X(s), Y(s) fitted to the waypoints resampled every 0.5 m (`fit_XY_s`
:60-72), curvature (x'y'' - y'x'') / (x'^2 + y'^2) of those cubics sampled every 0.25 m and fitted
by a cubic again (`compute_curvature_poly` :44-58), initial path heading atan2(Y'(0), X'(0))
(`get_reference_frenet` :76-86).  The Gazebo node feeds the solver (s, e_y, e_psi, v) =
(0, 0, -psi_start, speed) and K_coeffs (gazebo_sim_mpc_cmd_pub_frenet.jl:125-126).
"""
import numpy as np


def get_reference_frenet(path):
    """path: dict with 'x', 'y', 's' arrays (s increasing from the first waypoint).
    Returns (K_coeffs[4] highest degree first, psi_start, x_interp, y_interp) like :76-86."""
    x = np.asarray(path["x"], dtype=np.float64); y = np.asarray(path["y"], dtype=np.float64)
    s = np.asarray(path["s"], dtype=np.float64)
    s_fit = np.arange(s[0], s[-1], 0.5)                                     # :63
    x_coeffs = np.polyfit(s_fit, np.interp(s_fit, s, x), 3)                 # :65-70
    y_coeffs = np.polyfit(s_fit, np.interp(s_fit, s, y), 3)
    s_interp = np.arange(0.0, s[-1], 0.25)                                  # :79
    x_interp = np.polyval(x_coeffs, s_interp); y_interp = np.polyval(y_coeffs, s_interp)
    K_coeffs = curvature_poly(s_interp, x_coeffs, y_coeffs)
    dx0 = np.polyval(np.polyder(x_coeffs), 0.0); dy0 = np.polyval(np.polyder(y_coeffs), 0.0)
    return K_coeffs, float(np.arctan2(dy0, dx0)), x_interp, y_interp


def curvature_poly(s_interp, x_coeffs, y_coeffs):
    """compute_curvature_poly :44-58; x_coeffs / y_coeffs may carry a trailing batch axis (4, B)."""
    def d1(c, t):
        return c[2] + 2 * c[1] * t + 3 * c[0] * t ** 2
    def d2(c, t):
        return 2 * c[1] + 6 * c[0] * t
    t = s_interp if np.ndim(x_coeffs) == 1 else s_interp[:, None]
    dx, dy, ddx, ddy = d1(x_coeffs, t), d1(y_coeffs, t), d2(x_coeffs, t), d2(y_coeffs, t)
    K_meas = (dx * ddy - dy * ddx) / (dx * dx + dy * dy)
    return np.polyfit(s_interp, K_meas, 3)


_pinv_cache = {}


def _pinv(x):
    """Least-squares cubic fit as a fixed 4 x n matrix (what np.polyfit solves, without its per-call scaling)."""
    key = (x.size, float(x[0]), float(x[-1]))
    if key not in _pinv_cache:
        scale = np.sqrt((np.vander(x, 4) ** 2).sum(axis=0))
        _pinv_cache[key] = np.linalg.pinv(np.vander(x, 4) / scale) / scale[:, None]
    return _pinv_cache[key]


def fit_windows(xw, yw, s_end):
    """Batched form for synthetic workloads: B windows resampled on the SAME arc-length grid
    arange(0, s_end, 0.5) (xw, yw: (B, n_grid)).  The fits of get_reference_frenet with s[0] = 0,
    s[-1] = s_end, each problem on its own with fixed-shape products, so that a problem's result does
    not depend on which batch (or which GPU's slice) it is generated in.  Agrees with np.polyfit to
    rounding.  Returns K_coeffs (B, 4) and psi_start (B,)."""
    s_fit = np.arange(0.0, s_end, 0.5)
    assert xw.shape[1] == s_fit.size
    s_interp = np.arange(0.0, s_end, 0.25)
    P1, P2 = _pinv(s_fit), _pinv(s_interp)
    B = xw.shape[0]
    K = np.empty((B, 4)); psi0 = np.empty(B)
    for q in range(B):
        xc = P1.dot(xw[q]); yc = P1.dot(yw[q])
        dx = xc[2] + 2 * xc[1] * s_interp + 3 * xc[0] * s_interp ** 2
        dy = yc[2] + 2 * yc[1] * s_interp + 3 * yc[0] * s_interp ** 2
        ddx = 2 * xc[1] + 6 * xc[0] * s_interp; ddy = 2 * yc[1] + 6 * yc[0] * s_interp
        K[q] = P2.dot((dx * ddy - dy * ddx) / (dx * dx + dy * dy))
        psi0[q] = np.arctan2(yc[2], xc[2])
    return K, psi0
