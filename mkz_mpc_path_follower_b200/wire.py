"""ROS 1 wire format of the three messages on the hot path, without ROS.

    msg/state_est.msg   Header header; float64 x y psi v lat lon a df      (input of the MPC node)
    msg/MPC_cmd.msg     Header header; float64 accel_cmd steer_angle_cmd   (its output)
    msg/mpc_path.msg    Header header; float64[] xs ys psis                (target / predicted path)

ROS 1 serialisation is little-endian and field-ordered: std_msgs/Header = uint32 seq, time stamp
(uint32 secs, uint32 nsecs), string frame_id (uint32 length + bytes); float64 = 8 bytes;
float64[] = uint32 count + 8*count bytes.  These helpers let a bridge process feed
`mpcb200_solve_batch` from recorded or live `state_est` messages and publish `MPC_cmd` /
`mpc_path` without importing rospy; the batch variants go straight to/from the problem-major
arrays of include/mpc_b200.h.  Nothing here is on the timed path.
"""
import struct

import numpy as np

_HDR = struct.Struct("<III")


def pack_header(seq=0, secs=0, nsecs=0, frame_id=""):
    fid = frame_id.encode("utf-8")
    return _HDR.pack(seq & 0xFFFFFFFF, secs & 0xFFFFFFFF, nsecs & 0xFFFFFFFF) + struct.pack("<I", len(fid)) + fid


def unpack_header(buf, off=0):
    seq, secs, nsecs = _HDR.unpack_from(buf, off)
    (n,) = struct.unpack_from("<I", buf, off + 12)
    end = off + 16 + n
    if end > len(buf):
        raise ValueError("truncated Header")
    return {"seq": seq, "secs": secs, "nsecs": nsecs, "frame_id": bytes(buf[off + 16:end]).decode("utf-8")}, end


# ---- state_est (msg/state_est.msg:1-9) ----
STATE_EST_FIELDS = ("x", "y", "psi", "v", "lat", "lon", "a", "df")


def pack_state_est(x, y, psi, v, lat=0.0, lon=0.0, a=0.0, df=0.0, **hdr):
    return pack_header(**hdr) + struct.pack("<8d", x, y, psi, v, lat, lon, a, df)


def unpack_state_est(buf):
    h, off = unpack_header(buf)
    if len(buf) - off != 64:
        raise ValueError("state_est: expected 64 payload bytes, got %d" % (len(buf) - off))
    vals = struct.unpack_from("<8d", buf, off)
    return dict(zip(STATE_EST_FIELDS, vals), header=h)


def states_from_messages(msgs):
    """Serialized state_est messages (one per vehicle) -> state[B][4] = x, y, psi, v: exactly the four
    fields state_est_callback latches (mpc_cmd_pub.jl:72-84); a and df are not used by the node."""
    out = np.empty((len(msgs), 4))
    for i, m in enumerate(msgs):
        d = unpack_state_est(m)
        out[i] = (d["x"], d["y"], d["psi"], d["v"])
    return out


# ---- MPC_cmd (msg/MPC_cmd.msg:1-3) ----
def pack_mpc_cmd(accel_cmd, steer_angle_cmd, **hdr):
    return pack_header(**hdr) + struct.pack("<2d", accel_cmd, steer_angle_cmd)


def unpack_mpc_cmd(buf):
    h, off = unpack_header(buf)
    if len(buf) - off != 16:
        raise ValueError("MPC_cmd: expected 16 payload bytes, got %d" % (len(buf) - off))
    a, d = struct.unpack_from("<2d", buf, off)
    return {"accel_cmd": a, "steer_angle_cmd": d, "header": h}


def commands_to_messages(u0, stopped=None, secs=0, nsecs=0):
    """u0[B][2] = acc, d_f (solve_model order) -> serialized MPC_cmd per vehicle.  Vehicles whose
    stop latch is set get (-1.0, 0.0) like mpc_cmd_pub.jl:148-153."""
    u0 = np.asarray(u0, dtype=np.float64)
    msgs = []
    for i in range(u0.shape[0]):
        if stopped is not None and stopped[i]:
            msgs.append(pack_mpc_cmd(-1.0, 0.0, seq=i, secs=secs, nsecs=nsecs))
        else:
            msgs.append(pack_mpc_cmd(float(u0[i, 0]), float(u0[i, 1]), seq=i, secs=secs, nsecs=nsecs))
    return msgs


# ---- mpc_path (msg/mpc_path.msg:1-4) ----
def pack_mpc_path(xs, ys, psis, **hdr):
    out = [pack_header(**hdr)]
    for arr in (xs, ys, psis):
        a = np.ascontiguousarray(arr, dtype="<f8")
        out.append(struct.pack("<I", a.size))
        out.append(a.tobytes())
    return b"".join(out)


def unpack_mpc_path(buf):
    h, off = unpack_header(buf)
    arrs = []
    for _ in range(3):
        (n,) = struct.unpack_from("<I", buf, off)
        off += 4
        if off + 8 * n > len(buf):
            raise ValueError("mpc_path: truncated array")
        arrs.append(np.frombuffer(buf, dtype="<f8", count=n, offset=off).copy())
        off += 8 * n
    if off != len(buf):
        raise ValueError("mpc_path: %d trailing bytes" % (len(buf) - off))
    return {"xs": arrs[0], "ys": arrs[1], "psis": arrs[2], "header": h}


def predicted_path_message(traj_row, N, **hdr):
    """One row of `traj` (get_solver_results order x, y, v, psi, d_f, acc) -> the mpc_path message the
    node publishes from res[1], res[2], res[4] (mpc_cmd_pub.jl:143-147)."""
    t = np.asarray(traj_row, dtype=np.float64)
    return pack_mpc_path(t[0:N + 1], t[N + 1:2 * (N + 1)], t[3 * (N + 1):4 * (N + 1)], **hdr)


def reference_from_message(buf, N):
    """A target_path message (mpc_path with N+1 waypoints) -> ref[3][N+1] for mpcb200_solve_batch."""
    d = unpack_mpc_path(buf)
    if not (len(d["xs"]) == len(d["ys"]) == len(d["psis"]) == N + 1):
        raise ValueError("mpc_path: expected %d waypoints" % (N + 1))
    return np.stack((d["xs"], d["ys"], d["psis"]))
