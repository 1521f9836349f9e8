"""Synthetic open-loop batches of SURVEY.md section 8(d): perturbed states along the
recorded paths, previous command from the recording, time-mode references.

Counter-based and sliceable: problem j draws its 8 uniforms from Philox(key=KEY)
advanced by 2*j blocks, so any contiguous slice [b0, b1) of a batch is reproduced
exactly on any GPU count.
"""
import math
import numpy as np

from . import paths as _paths
from .gps_ref_traj import GPSRefTrajectory

KEY = 20261018

_traj_cache = {}


def _traj(path_id, N, dt):
    k = (path_id, N, dt)
    if k not in _traj_cache:
        _traj_cache[k] = GPSRefTrajectory(mat_filename=path_id, traj_horizon=N, traj_dt=dt)
    return _traj_cache[k]


def _uniforms(b0, b1):
    bg = np.random.Philox(key=KEY)
    bg.advance(2 * b0)
    raw = bg.random_raw(8 * (b1 - b0)).reshape(b1 - b0, 8)
    return ((raw >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def make_batch(B, N, path_ids=(1, 2, 3), dt=0.2, b0=0, v_des=1.0):
    """Problems b0 .. b0+B-1 of the infinite synthetic stream.  Returns dict with
    state (B,4) x,y,psi,v ; ref (B,3,N+1) ; u_prev (B,2) d_f_current, acc_current ;
    v_des (B,) ; path (B,) int ; idx (B,) sample index."""
    u = _uniforms(b0, b0 + B)
    j = np.arange(b0, b0 + B)
    pid = np.asarray(path_ids)[j % len(path_ids)]
    state = np.empty((B, 4))
    u_prev = np.empty((B, 2))
    ref = np.empty((B, 3, N + 1))
    idx = np.empty(B, dtype=np.int64)
    # Box-Muller normals from uniform pairs (1,2) and (3,4)
    r1 = np.sqrt(-2.0 * np.log(u[:, 1])); r2 = np.sqrt(-2.0 * np.log(u[:, 3]))
    n0 = r1 * np.cos(2 * math.pi * u[:, 2]); n1 = r1 * np.sin(2 * math.pi * u[:, 2])
    n2 = r2 * np.cos(2 * math.pi * u[:, 4]); n3 = r2 * np.sin(2 * math.pi * u[:, 4])
    for p in sorted(set(pid.tolist())):
        m = pid == p
        g = _traj(p, N, dt)
        data = _paths.load_path(p)
        n_p = g.trajectory.shape[0]
        span = n_p - int(math.ceil(N * dt / 0.01))  # keep the horizon inside the recording
        i = np.minimum((u[m, 0] * span).astype(np.int64), span - 1)
        idx[m] = i
        tr = g.trajectory
        state[m, 0] = tr[i, 4] + 0.3 * n0[m]
        state[m, 1] = tr[i, 5] + 0.3 * n1[m]
        state[m, 2] = tr[i, 3] + 0.05 * n2[m]
        state[m, 3] = np.clip(data['v'][i] + 0.5 * n3[m], 0.0, 20.0)
        u_prev[m, 0] = np.clip(data['df'][i], -0.5, 0.5)
        u_prev[m, 1] = np.clip(data['a'][i], -1.0, 1.0)
        r, _ = g.get_waypoints_batch(state[m, 0], state[m, 1], state[m, 2])
        ref[m] = r
    return {
        "state": np.ascontiguousarray(state), "ref": np.ascontiguousarray(ref),
        "u_prev": np.ascontiguousarray(u_prev), "v_des": np.full(B, float(v_des)),
        "path": pid.astype(np.int32), "idx": idx,
    }


def reference_start(batch, N):
    """A start point in get_solver_results order (x, y, v, psi, d_f, acc) that follows the reference
    waypoints at the measured speed with zero inputs -- what a caller without a previous solution
    would pass as `warm` for long horizons, where the all-zero `start=0.0` point is far away."""
    B = batch["state"].shape[0]
    w = np.zeros((B, 6 * N + 4))
    w[:, 0:N + 1] = batch["ref"][:, 0]
    w[:, N + 1:2 * (N + 1)] = batch["ref"][:, 1]
    w[:, 2 * (N + 1):3 * (N + 1)] = batch["state"][:, 3:4]
    w[:, 3 * (N + 1):4 * (N + 1)] = batch["ref"][:, 2]
    return w


def make_frenet_batch(B, N, path_ids=(1, 2, 3), dt=0.2, b0=0, window=40.0):
    """Synthetic batch for the Frenet-frame variant (MKZMPCPathFollowerFrenet.jl), same stream and
    perturbations as make_batch: problem j sits at a sample of a recorded path; the next `window`
    metres of the path, seen from the (perturbed) vehicle pose like the vehicle-frame nav_msgs/Path of
    the Gazebo node, are fitted by frenet_ref (K_coeffs, psi_start).  State = (s, e_y, e_psi, v) =
    (0, lateral offset of the perturbed pose, -psi_start, perturbed speed)
    (gazebo_sim_mpc_cmd_pub_frenet.jl:125 uses e_y = 0; the offset makes the problems non-trivial);
    v_des = the recorded speed + 1 m/s clipped to [1, 15] (C_ev = 0.5 is active in this variant).
    Returns dict with state (B,4), kpoly (B,4), u_prev (B,2), v_des (B,), path, idx."""
    from . import frenet_ref
    u = _uniforms(b0, b0 + B)
    j = np.arange(b0, b0 + B)
    pid = np.asarray(path_ids)[j % len(path_ids)]
    state = np.zeros((B, 4)); u_prev = np.empty((B, 2)); kpoly = np.empty((B, 4)); v_des = np.empty(B)
    idx = np.empty(B, dtype=np.int64)
    r1 = np.sqrt(-2.0 * np.log(u[:, 1])); r2 = np.sqrt(-2.0 * np.log(u[:, 3]))
    n0 = r1 * np.cos(2 * math.pi * u[:, 2]); n1 = r1 * np.sin(2 * math.pi * u[:, 2])
    n2 = r2 * np.cos(2 * math.pi * u[:, 4]); n3 = r2 * np.sin(2 * math.pi * u[:, 4])
    s_fit = np.arange(0.0, window, 0.5)
    for p in sorted(set(pid.tolist())):
        m = np.nonzero(pid == p)[0]
        g = _traj(p, N, dt)
        data = _paths.load_path(p)
        tr = g.trajectory
        s_all = tr[:, 6]
        last = int(np.searchsorted(s_all, s_all[-1] - window - 1.0))   # keep the window inside the recording
        i = np.minimum((u[m, 0] * last).astype(np.int64), last - 1)
        idx[m] = i
        X0 = tr[i, 4] + 0.3 * n0[m]; Y0 = tr[i, 5] + 0.3 * n1[m]; yaw = tr[i, 3] + 0.05 * n2[m]
        xw = np.empty((m.size, s_fit.size)); yw = np.empty_like(xw)
        for q in range(m.size):
            sq = s_all[i[q]] + s_fit
            dx = np.interp(sq, s_all, tr[:, 4]) - X0[q]; dy = np.interp(sq, s_all, tr[:, 5]) - Y0[q]
            c, s_ = math.cos(yaw[q]), math.sin(yaw[q])
            xw[q] = c * dx + s_ * dy; yw[q] = -s_ * dx + c * dy           # vehicle frame
        K, psi_start = frenet_ref.fit_windows(xw, yw, window)
        kpoly[m] = K
        # lateral offset of the vehicle (the origin) from the fitted path start, along the path normal
        state[m, 1] = -(-np.sin(psi_start) * xw[:, 0] + np.cos(psi_start) * yw[:, 0])
        state[m, 2] = -psi_start
        state[m, 3] = np.clip(data['v'][i] + 0.5 * n3[m], 0.0, 20.0)
        u_prev[m, 0] = np.clip(data['df'][i], -0.5, 0.5)
        u_prev[m, 1] = np.clip(data['a'][i], -1.0, 1.0)
        v_des[m] = np.clip(data['v'][i] + 1.0, 1.0, 15.0)
    return {"state": np.ascontiguousarray(state), "kpoly": np.ascontiguousarray(kpoly),
            "u_prev": np.ascontiguousarray(u_prev), "v_des": v_des, "path": pid.astype(np.int32), "idx": idx}
