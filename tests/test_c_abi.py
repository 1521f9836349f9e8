"""The C ABI seen from C: struct layouts pinned against the ctypes binding and the Julia shims, and a plain-C caller
(tests/c_abi/mpc_cmd_loop.c, the call sequence of julia/MKZMPCPathFollower.jl behind mpc_cmd_pub.jl:115-141) driving
libmpc_b200.so in its own process, compared with the Python mirror bit for bit."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CABI = os.path.join(ROOT, "tests", "c_abi")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def _build(name, link=False, out_dir=None):
    out = os.path.join(out_dir or CABI, name)
    cmd = [GCC, "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(CABI, name + ".c"), "-o", out]
    if link:
        pkg = os.path.join(ROOT, "mkz_mpc_path_follower_b200")
        cmd += ["-L", pkg, "-lmpc_b200", "-Wl,-rpath," + pkg]
    subprocess.check_call(cmd)
    return out


def _c_layout(tmp_path):
    exe = _build("layout", out_dir=str(tmp_path))
    lay = {}
    for line in subprocess.check_output([exe], text=True).splitlines():
        f = line.split()
        lay[f[0] + ("" if f[1] != "sizeof" else ".sizeof")] = tuple(int(x) for x in f[1:] if x.isdigit()) if f[1] != "sizeof" else int(f[2])
    return lay


def test_struct_layout_matches_ctypes_and_julia(tmp_path):
    from mkz_mpc_path_follower_b200 import capi
    lay = _c_layout(tmp_path)
    assert lay["mpcb200_config.sizeof"] == C.sizeof(capi.Config)
    assert lay["mpcb200_stats.sizeof"] == C.sizeof(capi.Stats)
    for cls, cname in ((capi.Config, "mpcb200_config"), (capi.Stats, "mpcb200_stats")):
        c_fields = [k.split(".")[1] for k in lay if k.startswith(cname + ".") and not k.endswith("sizeof")]
        assert c_fields == [n for n, _ in cls._fields_], cname            # same fields, same order
        for n, _ in cls._fields_:
            d = getattr(cls, n)
            assert (d.offset, d.size) == lay["%s.%s" % (cname, n)], (cname, n)
    assert lay["MPCB200_VERSION"] == (capi.lib().mpcb200_version(),)
    # the Julia shims' `Config`: same field names in the same order, with types of the same size (Julia lays a
    # struct of isbits fields out like C)
    jl_size = {"Int32": 4, "Float64": 8, "NTuple{8,Int32}": 32}
    for fn in ("MKZMPCPathFollower.jl", "MKZMPCPathFollowerFrenet.jl"):
        src = open(os.path.join(ROOT, "julia", fn)).read()
        body = re.search(r"mutable struct Config[^\n]*\n(.*?)Config\(\) = new\(\)", src, re.S).group(1)
        fields = re.findall(r"(\w+)::([\w{},]+)", body)
        assert [n for n, _ in fields] == [n for n, _ in capi.Config._fields_], fn
        off = 0
        for n, t in fields:
            sz = jl_size[t]
            al = 4 if t != "Float64" else 8
            off = (off + al - 1) // al * al
            assert (off, sz) == lay["mpcb200_config." + n], (fn, n)
            off += sz
        assert (off + 7) // 8 * 8 == lay["mpcb200_config.sizeof"]
        # every ccall of the shim names an exported symbol
        for sym in set(re.findall(r"\(:(mpcb200_\w+), libmpc\)", src)):
            assert hasattr(capi.lib(), sym), (fn, sym)


def test_c_caller_compiles_and_links(tmp_path):
    exe = _build("mpc_cmd_loop", link=True, out_dir=str(tmp_path))
    assert os.path.exists(exe)


def _fixture(path, N, T):
    """State / reference sequence of a closed-loop run on path 1 (host plant + host generator), N = 8."""
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    from mkz_mpc_path_follower_b200.mpc_path_follower import MKZMPCPathFollower
    from mkz_mpc_path_follower_b200.vehicle_simulator import VehicleSimulator
    g = GPSRefTrajectory(mat_filename=1, traj_horizon=N, traj_dt=0.2)
    sim = VehicleSimulator(X0=g.trajectory[200, 4] + 0.4, Y0=g.trajectory[200, 5] - 0.3, Psi0=g.trajectory[200, 3] + 0.05)
    kmpc = MKZMPCPathFollower(N=N)                 # module load (initial solve)
    load = kmpc._last
    kmpc.update_cost(9.0, 9.0, 10.0, 0.0, 100.0, 1000.0, 0.0, 0.0)
    rows, outs = [], []
    for t in range(T):
        for _ in range(10):
            sim.update_vehicle_model()
        x, y, psi, v = sim.state_est()[:4]
        xr, yr, pr, stop = g.get_waypoints(x, y, psi)
        kmpc.update_init_cond(x, y, psi, v)
        kmpc.update_reference(xr, yr, pr, 1.0)
        a_opt, df_opt, is_opt = kmpc.solve_model()
        kmpc.update_current_input(df_opt, a_opt)
        sim.mpc_cmd(a_opt, df_opt)
        rows.append(np.concatenate(([x, y, psi, v], xr, yr, pr, [1.0])))
        outs.append((int(kmpc._last["status"][0]), int(kmpc._last["iters"][0]), a_opt, df_opt, float(kmpc._last["cost"][0])))
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", N, T))
        f.write(np.asarray(rows, dtype="<f8").tobytes())
    return load, outs


@pytest.mark.gpu
def test_c_caller_drives_the_library_like_the_python_mirror(tmp_path):
    """20 closed-loop steps: a C process (default_config -> create -> module-load solve -> set_cost -> per step
    solve_batch(B = 1) with warm in/out and command feedback) prints exactly the commands, costs, statuses and iteration
    counts the Python mirror got on the same inputs."""
    N, T = 8, 20
    fx = str(tmp_path / "fixture.bin")
    load, outs = _fixture(fx, N, T)
    exe = _build("mpc_cmd_loop", link=True, out_dir=str(tmp_path))
    lines = subprocess.check_output([exe, fx], text=True).splitlines()
    assert lines[0].split()[0] == "load"
    f = lines[0].split()
    assert (int(f[1]), int(f[2])) == (int(load["status"][0]), int(load["iters"][0]))
    assert float.fromhex(f[3]) == load["u0"][0, 0] and float.fromhex(f[4]) == load["u0"][0, 1]
    assert len(lines) == T + 2 and lines[-1].startswith("stats 1 ")
    for t in range(T):
        f = lines[1 + t].split()
        st, it, a, d, c = outs[t]
        assert (int(f[0]), int(f[1]), int(f[2])) == (t, st, it), (t, f)
        assert float.fromhex(f[3]) == a and float.fromhex(f[4]) == d and float.fromhex(f[5]) == c, (t, f)
    assert all(o[0] == 0 for o in outs)
