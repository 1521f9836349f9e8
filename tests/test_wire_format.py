"""ROS 1 wire format of state_est / MPC_cmd / mpc_path (msg/*.msg of the reference), checked against
byte strings written out by hand from the ROS serialisation rules (little-endian, uint32 seq,
uint32 secs, uint32 nsecs, uint32-length-prefixed frame_id, float64 fields, uint32-count arrays)."""
import struct

import numpy as np
import pytest

from mkz_mpc_path_follower_b200 import wire


def test_state_est_golden_bytes():
    # seq=7, stamp=(12, 500), frame_id="map", then x y psi v lat lon a df
    golden = (b"\x07\x00\x00\x00" b"\x0c\x00\x00\x00" b"\xf4\x01\x00\x00" b"\x03\x00\x00\x00map" +
              struct.pack("<8d", 1.5, -2.25, 0.125, 8.0, 37.917929, -122.331798, 0.5, -0.01))
    assert wire.pack_state_est(1.5, -2.25, 0.125, 8.0, 37.917929, -122.331798, 0.5, -0.01,
                               seq=7, secs=12, nsecs=500, frame_id="map") == golden
    d = wire.unpack_state_est(golden)
    assert d["header"] == {"seq": 7, "secs": 12, "nsecs": 500, "frame_id": "map"}
    assert (d["x"], d["y"], d["psi"], d["v"], d["a"], d["df"]) == (1.5, -2.25, 0.125, 8.0, 0.5, -0.01)
    assert len(golden) == 16 + 3 + 64
    with pytest.raises(ValueError):
        wire.unpack_state_est(golden[:-8])


def test_mpc_cmd_golden_bytes_and_stop_latch():
    golden = b"\x00" * 12 + b"\x00\x00\x00\x00" + struct.pack("<2d", -0.75, 0.0625)
    assert wire.pack_mpc_cmd(-0.75, 0.0625) == golden
    assert wire.unpack_mpc_cmd(golden)["steer_angle_cmd"] == 0.0625
    msgs = wire.commands_to_messages(np.array([[0.3, -0.1], [0.9, 0.2]]), stopped=[False, True])
    assert wire.unpack_mpc_cmd(msgs[0])["accel_cmd"] == 0.3
    m1 = wire.unpack_mpc_cmd(msgs[1])     # mpc_cmd_pub.jl:148-153
    assert (m1["accel_cmd"], m1["steer_angle_cmd"]) == (-1.0, 0.0)


def test_mpc_path_roundtrip_and_layouts():
    N = 8
    xs, ys, ps = np.arange(N + 1) * 1.0, np.arange(N + 1) * -2.0, np.linspace(-3, 3, N + 1)
    buf = wire.pack_mpc_path(xs, ys, ps, seq=1, frame_id="")
    assert len(buf) == 16 + 3 * (4 + 8 * (N + 1))
    assert buf[16:20] == struct.pack("<I", N + 1) and buf[20:28] == struct.pack("<d", 0.0)
    ref = wire.reference_from_message(buf, N)
    assert ref.shape == (3, N + 1) and (ref[0] == xs).all() and (ref[2] == ps).all()
    # a traj row in get_solver_results order x, y, v, psi, d_f, acc -> message publishes x, y, psi
    traj = np.concatenate((xs, ys, np.full(N + 1, 5.0), ps, np.zeros(N), np.zeros(N)))
    d = wire.unpack_mpc_path(wire.predicted_path_message(traj, N))
    assert (d["xs"] == xs).all() and (d["ys"] == ys).all() and (d["psis"] == ps).all()
    with pytest.raises(ValueError):
        wire.reference_from_message(buf, N + 1)


def test_states_from_messages_takes_only_xypsiv():
    msgs = [wire.pack_state_est(i, i + 0.5, 0.1 * i, 2.0 * i, a=99.0, df=99.0) for i in range(4)]
    st = wire.states_from_messages(msgs)
    assert st.shape == (4, 4) and (st[:, 3] == [0, 2, 4, 6]).all() and (st[:, 1] == [0.5, 1.5, 2.5, 3.5]).all()


def test_roundtrip_properties():
    """Property check (hypothesis): pack -> unpack is the identity for arbitrary finite payloads and headers."""
    from hypothesis import given, settings, strategies as st
    f = st.floats(allow_nan=False, allow_infinity=False, width=64)

    @settings(max_examples=60, deadline=None)
    @given(st.lists(f, min_size=8, max_size=8), st.integers(0, 2**32 - 1), st.integers(0, 2**32 - 1), st.text(max_size=12))
    def state(vals, seq, secs, frame):
        d = wire.unpack_state_est(wire.pack_state_est(*vals, seq=seq, secs=secs, nsecs=7, frame_id=frame))
        assert [d[k] for k in wire.STATE_EST_FIELDS] == vals
        assert d["header"] == {"seq": seq, "secs": secs, "nsecs": 7, "frame_id": frame}

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 40).flatmap(lambda n: st.tuples(*[st.lists(f, min_size=n, max_size=n)] * 3)))
    def path(arrs):
        d = wire.unpack_mpc_path(wire.pack_mpc_path(*arrs))
        assert all((d[k] == np.array(a)).all() for k, a in zip(("xs", "ys", "psis"), arrs))

    state()
    path()
