"""TEST ONLY: ctypes binding of the CPU warp emulation of the CUDA solver (tests/emu)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libmpc_emu.so")
_SRCS = [os.path.join(_HERE, "emu_driver.cpp"), os.path.join(_HERE, "warp_emu.h"),
         os.path.join(_ROOT, "mkz_mpc_path_follower_b200", "csrc", "mpc_kernel.cuh"),
         os.path.join(_ROOT, "mkz_mpc_path_follower_b200", "csrc", "tpp_solver.cuh")]


class KCfg(C.Structure):
    _fields_ = [("N", C.c_int), ("max_iter", C.c_int), ("start_mode", C.c_int), ("pad_", C.c_int),
                ("dt", C.c_double), ("dtc", C.c_double), ("La", C.c_double), ("Lb", C.c_double),
                ("vmin", C.c_double), ("vmax", C.c_double), ("amax", C.c_double), ("smax", C.c_double),
                ("admax", C.c_double), ("sdmax", C.c_double), ("tol", C.c_double), ("w", C.c_double * 8),
                # derived fields, filled by kcfg_finalize() inside the driver
                ("vLo", C.c_double), ("vHi", C.c_double), ("aLo", C.c_double), ("aHi", C.c_double),
                ("dLo", C.c_double), ("dHi", C.c_double), ("rHiFirst", C.c_double * 2), ("rHiLater", C.c_double * 2),
                ("rfrac", C.c_double), ("mu_min", C.c_double), ("dtLb", C.c_double),
                ("roles", C.c_void_p)]   # filled inside the driver


def build():
    if (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in _SRCS):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DMPC_HOST_EMU",
                               "-I" + _HERE, "-I" + os.path.join(_ROOT, "mkz_mpc_path_follower_b200", "csrc"),
                               "-x", "c++", _SRCS[0], "-o", _LIB])
    return _LIB


def kcfg_from_oracle(ocfg, start_mode=0):
    k = KCfg()
    k.N = ocfg.N; k.max_iter = ocfg.max_iter; k.start_mode = start_mode
    k.dt = ocfg.dt; k.dtc = ocfg.dt_control; k.La = ocfg.L_a; k.Lb = ocfg.L_b
    k.vmin = ocfg.v_min; k.vmax = ocfg.v_max; k.amax = ocfg.a_max; k.smax = ocfg.steer_max
    k.admax = ocfg.a_dmax; k.sdmax = ocfg.steer_dmax; k.tol = ocfg.tol
    for i in range(8):
        k.w[i] = ocfg.w[i]
    return k


def race_check(on=True):
    """Switch the emulator's shared-memory race check on (and reset its counter) or off."""
    C.CDLL(build()).emu_race_enable(1 if on else 0)


def race_count():
    lib = C.CDLL(build())
    lib.emu_race_count.restype = C.c_long
    return lib.emu_race_count()


def solve_batch(kcfg, state, ref, v_des, u_prev, warm=None, want_traj=False):
    lib = C.CDLL(build())
    assert lib.emu_kcfg_size() == C.sizeof(KCfg)
    dp = C.POINTER(C.c_double)
    N = kcfg.N
    B = state.shape[0]
    p = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    state = np.ascontiguousarray(state, dtype=np.float64); ref = np.ascontiguousarray(ref, dtype=np.float64)
    u_prev = np.ascontiguousarray(u_prev, dtype=np.float64)
    v_des = None if v_des is None else np.ascontiguousarray(v_des, dtype=np.float64)
    u0 = np.empty((B, 2)); cost = np.empty(B); status = np.empty(B, dtype=np.int32); iters = np.empty(B, dtype=np.int32)
    traj = np.empty((B, 6 * N + 4)) if want_traj else None
    lib.emu_solve_batch.argtypes = [C.POINTER(KCfg), C.c_long, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp]
    lib.emu_solve_batch(C.byref(kcfg), B, p(state), p(ref), p(v_des), p(u_prev),
                        None if warm is None else warm.ctypes.data_as(dp), p(u0), p(cost),
                        status.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)),
                        None if traj is None else traj.ctypes.data_as(dp))
    return {"u0": u0, "cost": cost, "status": status, "iters": iters, "traj": traj}


def solve_batch_tpp(kcfg, state, ref, v_des, u_prev, warm=None, want_traj=False, slots=3):
    """The thread-per-problem solver (csrc/tpp_solver.cuh) compiled for the host; problems round-robin over `slots` slots."""
    lib = C.CDLL(build())
    assert lib.emu_kcfg_size() == C.sizeof(KCfg)
    dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)
    N = kcfg.N
    B = state.shape[0]
    p = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    state = np.ascontiguousarray(state, dtype=np.float64); ref = np.ascontiguousarray(ref, dtype=np.float64)
    u_prev = np.ascontiguousarray(u_prev, dtype=np.float64)
    v_des = None if v_des is None else np.ascontiguousarray(v_des, dtype=np.float64)
    u0 = np.empty((B, 2)); cost = np.empty(B); status = np.empty(B, dtype=np.int32); iters = np.empty(B, dtype=np.int32)
    resto = np.empty(B, dtype=np.int32); ticks = np.empty(B, dtype=np.int64)
    traj = np.empty((B, 6 * N + 4)) if want_traj else None
    lib.emu_solve_batch_tpp.argtypes = [C.POINTER(KCfg), C.c_long, dp, dp, dp, dp, dp, dp, dp, ip, ip, dp, ip, C.c_long, C.POINTER(C.c_long)]
    lib.emu_solve_batch_tpp(C.byref(kcfg), B, p(state), p(ref), p(v_des), p(u_prev),
                            None if warm is None else warm.ctypes.data_as(dp), p(u0), p(cost),
                            status.ctypes.data_as(ip), iters.ctypes.data_as(ip),
                            None if traj is None else traj.ctypes.data_as(dp), resto.ctypes.data_as(ip), slots,
                            ticks.ctypes.data_as(C.POINTER(C.c_long)))
    return {"u0": u0, "cost": cost, "status": status, "iters": iters, "traj": traj, "n_resto": resto, "ticks": ticks}


def solve_batch_frenet(kcfg, state, kpoly, v_des, u_prev, warm=None, want_traj=False, tpp=False, slots=3):
    """Frenet-frame variant on the emulator; tpp: the thread-per-problem solver (TppSolverT<1>) instead of the warp kernel."""
    lib = C.CDLL(build())
    assert lib.emu_kcfg_size() == C.sizeof(KCfg)
    B = state.shape[0]
    N = kcfg.N
    dp = C.POINTER(C.c_double)
    p = lambda a: None if a is None else a.ctypes.data_as(dp)
    state = np.ascontiguousarray(state, dtype=np.float64); kpoly = np.ascontiguousarray(kpoly, dtype=np.float64)
    u_prev = np.ascontiguousarray(u_prev, dtype=np.float64)
    v_des = None if v_des is None else np.ascontiguousarray(v_des, dtype=np.float64)
    u0 = np.empty((B, 2)); cost = np.empty(B); status = np.empty(B, dtype=np.int32); iters = np.empty(B, dtype=np.int32)
    traj = np.empty((B, 6 * N + 4)) if want_traj else None
    if tpp:
        resto = np.empty(B, dtype=np.int32)
        lib.emu_solve_batch_tpp_frenet.argtypes = [C.POINTER(KCfg), C.c_long, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp,
                                                   C.POINTER(C.c_int), C.c_long]
        rc = lib.emu_solve_batch_tpp_frenet(C.byref(kcfg), B, p(state), p(kpoly), p(v_des), p(u_prev), p(warm), p(u0), p(cost),
                                            status.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)), p(traj),
                                            resto.ctypes.data_as(C.POINTER(C.c_int)), slots)
        assert rc == 0
        return {"u0": u0, "cost": cost, "status": status, "iters": iters, "traj": traj, "n_resto": resto}
    lib.emu_solve_batch_frenet.argtypes = [C.POINTER(KCfg), C.c_long, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp]
    rc = lib.emu_solve_batch_frenet(C.byref(kcfg), B, p(state), p(kpoly), p(v_des), p(u_prev), p(warm), p(u0), p(cost),
                                    status.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)), p(traj))
    assert rc == 0
    return {"u0": u0, "cost": cost, "status": status, "iters": iters, "traj": traj}


def rollout(kcfg, traj_table, pose0, T, track_using_time=True, target_vel=1.0, warm0=None, tpp=False):
    """All vehicles on the path whose (n,7) table is given.  warm0 (6N+4,): start point of the first solve (None = zeros).
    Horizons above 31 run one emulated block of 2-3 warps per vehicle.  tpp: the plant / waypoints / thread-per-problem-solve
    pipeline mpcb200_rollout uses for large fleets instead of the fused kernel.  Returns log (T,B,8), final (B,8)."""
    lib = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    pose0 = np.ascontiguousarray(np.atleast_2d(pose0), dtype=np.float64)
    B = pose0.shape[0]
    cols = [np.ascontiguousarray(traj_table[:, i]) for i in (0, 4, 5, 3, 6)]
    path_of = np.zeros(B, dtype=np.int32)
    log = np.zeros((T, B, 8)); final = np.zeros((B, 8))
    fn = lib.emu_rollout_tpp if tpp else lib.emu_rollout
    fn.argtypes = [C.POINTER(KCfg), C.c_long, C.c_int, dp, C.POINTER(C.c_int), C.c_int, dp, dp, dp, dp, dp,
                   C.c_int, C.c_double, dp, dp, dp]
    w0 = None if warm0 is None else np.ascontiguousarray(warm0, dtype=np.float64)
    fn(C.byref(kcfg), B, T, pose0.ctypes.data_as(dp), path_of.ctypes.data_as(C.POINTER(C.c_int)), traj_table.shape[0],
                    *[c.ctypes.data_as(dp) for c in cols], int(track_using_time), float(target_vel),
                    log.ctypes.data_as(dp), final.ctypes.data_as(dp), None if w0 is None else w0.ctypes.data_as(dp))
    return log, final


def solve_batch_on_path(kcfg, traj_table, state, u_prev, track_using_time=True, target_vel=1.0, path_of=None):
    """mpcb200_solve_batch_on_path on the emulator: waypoints generated by the kernel source from one path table (id 0).
    Returns dict with u0, cost, status, iters, traj, ref (B,3,N+1), stop."""
    lib = C.CDLL(build())
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    N = kcfg.N
    state = np.ascontiguousarray(state, dtype=np.float64); u_prev = np.ascontiguousarray(u_prev, dtype=np.float64)
    B = state.shape[0]
    cols = [np.ascontiguousarray(traj_table[:, i]) for i in (0, 4, 5, 3, 6)]
    path_of = np.zeros(B, dtype=np.int32) if path_of is None else np.ascontiguousarray(path_of, dtype=np.int32)
    u0 = np.zeros((B, 2)); cost = np.zeros(B); status = np.zeros(B, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
    traj = np.zeros((B, 6 * N + 4)); ref = np.zeros((B, 3, N + 1)); stop = np.zeros(B, dtype=np.int32)
    lib.emu_solve_batch_on_path.argtypes = [C.POINTER(KCfg), C.c_long, dp, ip, C.c_int, dp, dp, dp, dp, dp, C.c_int, C.c_double,
                                            dp, dp, dp, ip, ip, dp, dp, ip]
    lib.emu_solve_batch_on_path(C.byref(kcfg), B, state.ctypes.data_as(dp), path_of.ctypes.data_as(ip), traj_table.shape[0],
                                *[c.ctypes.data_as(dp) for c in cols], int(track_using_time), float(target_vel),
                                u_prev.ctypes.data_as(dp), u0.ctypes.data_as(dp), cost.ctypes.data_as(dp), status.ctypes.data_as(ip),
                                iters.ctypes.data_as(ip), traj.ctypes.data_as(dp), ref.ctypes.data_as(dp), stop.ctypes.data_as(ip))
    return {"u0": u0, "cost": cost, "status": status, "iters": iters, "traj": traj, "ref": ref, "stop": stop}


def rollout_frenet(kcfg, traj_table, pose0, T, window=40.0, target_vel=8.0, ey_from_path=True):
    """mpcb200_rollout_frenet on the emulator: all vehicles on the path whose (n,7) table is given.  Returns log (T,B,8), final (B,8)."""
    lib = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    pose0 = np.ascontiguousarray(np.atleast_2d(pose0), dtype=np.float64)
    B = pose0.shape[0]
    cols = [np.ascontiguousarray(traj_table[:, i]) for i in (0, 4, 5, 3, 6)]
    path_of = np.zeros(B, dtype=np.int32)
    log = np.zeros((T, B, 8)); final = np.zeros((B, 8))
    lib.emu_rollout_frenet.argtypes = [C.POINTER(KCfg), C.c_long, C.c_int, dp, C.POINTER(C.c_int), C.c_int, dp, dp, dp, dp, dp,
                                       C.c_double, C.c_double, C.c_int, dp, dp]
    lib.emu_rollout_frenet(C.byref(kcfg), B, T, pose0.ctypes.data_as(dp), path_of.ctypes.data_as(C.POINTER(C.c_int)), traj_table.shape[0],
                           *[c.ctypes.data_as(dp) for c in cols], float(window), float(target_vel), int(bool(ey_from_path)),
                           log.ctypes.data_as(dp), final.ctypes.data_as(dp))
    return log, final
