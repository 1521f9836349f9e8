"""Reference-trajectory generator: rospy-free restatement of the reference's
scripts/gps_utils/ref_gps_traj.py (class GPSRefTrajectory, :54-218).

Same class name, constructor keywords and `get_waypoints` contract; the three
rosparams the original reads (`lat0`, `lon0`, `yaw0`, ref_gps_traj.py:61-73) become
keyword arguments with the launch-file values as defaults
(launch/sim_path_follow.launch:18-20).  `get_waypoints_batch` is the vectorised
form used by the batched harness.
"""
import math
import numpy as np

from . import paths as _paths


def latlon_to_XY(lat0, lon0, lat1, lon1):
    """Equirectangular projection, ref_gps_traj.py:33-52.  Accepts scalars or arrays."""
    R_earth = 6371000  # meters
    delta_lat = np.radians(lat1 - lat0)
    delta_lon = np.radians(lon1 - lon0)
    lat_avg = 0.5 * (np.radians(lat1) + np.radians(lat0))
    X = R_earth * delta_lon * np.cos(lat_avg)
    Y = R_earth * delta_lat
    return X, Y


class GPSRefTrajectory(object):
    def __init__(self, mat_filename=None, traj_horizon=8, traj_dt=0.2,
                 lat0=_paths.LAT0, lon0=_paths.LON0, yaw0=_paths.YAW0, path_data=None):
        if mat_filename is None and path_data is None:
            raise ValueError('Invalid matfile specified.')  # ref_gps_traj.py:67-68
        self.traj_horizon = traj_horizon
        self.traj_dt = traj_dt
        if path_data is None:
            if str(mat_filename).endswith('.mat'):
                import scipy.io as sio
                m = sio.loadmat(mat_filename)
                path_data = {k: np.ravel(m[k]) for k in ('t', 'lat', 'lon', 'psi')}
            else:
                path_data = _paths.load_path(mat_filename)
        tms = np.asarray(path_data['t'], dtype=np.float64)
        lats = np.asarray(path_data['lat'], dtype=np.float64)
        lons = np.asarray(path_data['lon'], dtype=np.float64)
        yaws = np.asarray(path_data['psi'], dtype=np.float64)
        # ref_gps_traj.py:94-103 -- per-sample projection with math.* and a running sum of
        # segment lengths; done in a Python loop so that every rounding matches.
        Xs = np.empty_like(tms)
        Ys = np.empty_like(tms)
        cd = np.empty_like(tms)
        lat0r = math.radians(lat0)
        for i in range(len(lats)):
            dlat = math.radians(lats[i] - lat0)
            dlon = math.radians(lons[i] - lon0)
            lat_avg = 0.5 * (math.radians(lats[i]) + lat0r)
            X = 6371000 * dlon * math.cos(lat_avg)
            Y = 6371000 * dlat
            if i == 0:
                cd[i] = 0.0
            else:
                cd[i] = math.sqrt((X - Xs[i - 1]) ** 2 + (Y - Ys[i - 1]) ** 2) + cd[i - 1]
            Xs[i] = X
            Ys[i] = Y
        self.trajectory = np.column_stack((tms, lats, lons, yaws, Xs, Ys, cd))  # :106
        self.x_interp = None
        self.y_interp = None
        self.psi_interp = None

    def get_global_trajectory_reference(self):
        return self.trajectory

    def get_Xs(self):
        return self.trajectory[:, 4]

    def get_Ys(self):
        return self.trajectory[:, 5]

    def get_psis(self):
        return self.trajectory[:, 3]

    # ---- single query, ref_gps_traj.py:131-142 ----
    def get_waypoints(self, X_init, Y_init, yaw_init, v_target=None):
        XY_traj = self.trajectory[:, 4:6]
        xy_query = np.array([[X_init, Y_init]])
        diff_dists = np.sum((XY_traj - xy_query) ** 2, axis=1)
        closest_traj_ind = int(np.argmin(diff_dists))
        if v_target is not None:
            return self._waypoints(closest_traj_ind, yaw_init, v_target)
        return self._waypoints(closest_traj_ind, yaw_init, None)

    def _waypoints(self, ind, yaw_init, v_target):
        tr = self.trajectory
        if v_target is not None:  # distance mode, :172-186
            start = tr[ind, 6]
            q = [x * self.traj_dt * v_target + start for x in range(1, self.traj_horizon + 2)]
            absc = tr[:, 6]
        else:  # time mode, :188-202
            start = tr[ind, 0]
            q = [h * self.traj_dt + start for h in range(0, self.traj_horizon + 1)]
            absc = tr[:, 0]
        self.x_interp = np.interp(q, absc, tr[:, 4])
        self.y_interp = np.interp(q, absc, tr[:, 5])
        psi_ref = np.interp(q, absc, tr[:, 3])
        self.psi_interp = self._fix_heading_wraparound(psi_ref, yaw_init)
        stop_cmd = bool(self.x_interp[-1] == tr[-1, 4] and self.y_interp[-1] == tr[-1, 5])
        return self.x_interp, self.y_interp, self.psi_interp, stop_cmd

    @staticmethod
    def _fix_heading_wraparound(psi_ref, psi_current):  # :204-218
        check_1 = np.max(np.fabs(np.diff(psi_ref))) < np.pi
        check_2 = np.max(np.fabs(psi_ref - psi_current)) < np.pi
        if check_1 and check_2:
            return psi_ref
        for i in range(len(psi_ref)):
            p = psi_ref[i]
            cands = np.array([p, p + 2 * np.pi, p - 2 * np.pi])
            psi_ref[i] = cands[np.argmin(np.fabs(cands - psi_current))]
        return psi_ref

    # ---- batched query (host-side vectorisation of the same arithmetic) ----
    def get_waypoints_batch(self, X, Y, yaw, v_target=None, closest_ind=None):
        """X, Y, yaw: (B,) arrays.  Returns ref (B, 3, N+1) = x_ref, y_ref, psi_ref and
        stop (B,) bool.  `closest_ind` lets a caller that already knows the nearest sample
        skip the O(n) search."""
        tr = self.trajectory
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        yaw = np.asarray(yaw, dtype=np.float64)
        B = X.shape[0]
        H = self.traj_horizon
        if closest_ind is None:
            closest_ind = np.empty(B, dtype=np.int64)
            step = max(1, (1 << 22) // tr.shape[0])
            for b0 in range(0, B, step):
                sl = slice(b0, min(B, b0 + step))
                d = (tr[None, :, 4] - X[sl, None]) ** 2 + (tr[None, :, 5] - Y[sl, None]) ** 2
                closest_ind[sl] = np.argmin(d, axis=1)
        if v_target is not None:
            absc = tr[:, 6]
            k = np.arange(1, H + 2, dtype=np.float64)
            q = k[None, :] * self.traj_dt * v_target + absc[closest_ind][:, None]
        else:
            absc = tr[:, 0]
            k = np.arange(0, H + 1, dtype=np.float64)
            q = k[None, :] * self.traj_dt + absc[closest_ind][:, None]
        ref = np.empty((B, 3, H + 1), dtype=np.float64)
        ref[:, 0, :] = np.interp(q.ravel(), absc, tr[:, 4]).reshape(B, H + 1)
        ref[:, 1, :] = np.interp(q.ravel(), absc, tr[:, 5]).reshape(B, H + 1)
        psi = np.interp(q.ravel(), absc, tr[:, 3]).reshape(B, H + 1)
        chk1 = np.max(np.fabs(np.diff(psi, axis=1)), axis=1) < np.pi
        chk2 = np.max(np.fabs(psi - yaw[:, None]), axis=1) < np.pi
        fix = ~(chk1 & chk2)
        if np.any(fix):
            p = psi[fix]
            cands = np.stack((p, p + 2 * np.pi, p - 2 * np.pi), axis=-1)
            best = np.argmin(np.fabs(cands - yaw[fix][:, None, None]), axis=-1)
            psi[fix] = np.take_along_axis(cands, best[..., None], axis=-1)[..., 0]
        ref[:, 2, :] = psi
        stop = (ref[:, 0, -1] == tr[-1, 4]) & (ref[:, 1, -1] == tr[-1, 5])
        return ref, stop
