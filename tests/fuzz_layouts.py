#!/usr/bin/env python
"""Fuzz of the two device layouts' SOURCES on the CPU (tests/emu: csrc/mpc_kernel.cuh and csrc/tpp_solver.cuh compiled with g++)
against the oracle, on inputs the synthetic workload does not produce: larger perturbations, random weights, speeds at and
outside the bounds, previous commands outside the box, stand-still, warm starts from another problem's solution; for the
Frenet-frame variant also curvatures with 1 - e_y K(s) <= 0 (outside the model's domain).
    python tests/fuzz_layouts.py [seconds] [seed] [frenet | rollout]
Prints every problem where the three disagree (status, or |du| > 1e-5 with all three Optimal).  Not collected by pytest (a
long-running hunt, not a check); tests/test_tpp_emu.py::test_fuzz_classes runs one small round of every input class.
What it found is in DESIGN.md 4."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "emu"))

XY_KINDS, FRENET_KINDS = 7, 5


def perturb_xy(rng, kind, b, N):
    """Input class `kind` applied to a workload.make_batch batch: (state, ref, v_des, u_prev, weights or None)."""
    B = b["state"].shape[0]
    st, up, vd, ref = b["state"].copy(), b["u_prev"].copy(), b["v_des"].copy(), b["ref"].copy()
    weights = None
    if kind == 1:      # large pose errors
        st[:, 0] += rng.normal(0, 3.0, B); st[:, 1] += rng.normal(0, 3.0, B); st[:, 2] += rng.normal(0, 0.6, B)
    elif kind == 2:    # speeds at / outside the bounds, desired speeds anywhere
        st[:, 3] = rng.choice([0.0, 1e-9, 0.5, 19.999, 20.0, 20.0 + 1e-9, 25.0, -0.1], B)
        vd = rng.uniform(-2.0, 25.0, B)
    elif kind == 3:    # previous commands at / outside the input box
        up[:, 0] = rng.choice([-0.7, -0.5, -0.5 + 1e-12, 0.0, 0.5, 0.56, 0.62], B)
        up[:, 1] = rng.choice([-3.5, -3.0, -1.0, 0.0, 2.0, 2.3, 3.0], B)
    elif kind == 4:    # random weights, some zero
        weights = [float(w) for w in rng.choice([0.0, 1e-3, 1.0, 9.0, 100.0, 1e4], 8)]
    elif kind == 5:    # stand-still at a far reference
        st[:, 3] = 0.0; ref[:, 0, :] += 40.0
    elif kind == 6:    # a reference that jumps
        ref[:, :2, N // 2:] += rng.normal(0, 5.0, (B, 2, 1))
    return st, ref, vd, up, weights


def perturb_frenet(rng, kind, b, N):
    """... to a workload.make_frenet_batch batch: (state, kpoly, v_des, u_prev, weights or None)."""
    B = b["state"].shape[0]
    st, up, vd, kp = b["state"].copy(), b["u_prev"].copy(), b["v_des"].copy(), b["kpoly"].copy()
    weights = None
    if kind == 1:      # large lateral / heading errors
        st[:, 1] += rng.normal(0, 2.0, B); st[:, 2] += rng.normal(0, 0.5, B)
    elif kind == 2:    # speeds at / outside the bounds, previous commands outside the box
        st[:, 3] = rng.choice([0.0, 1e-9, 0.5, 19.999, 20.0, 25.0, -0.1], B)
        up[:, 0] = rng.choice([-0.7, -0.5, 0.0, 0.5, 0.62], B)
        up[:, 1] = rng.choice([-3.5, -1.0, 0.0, 2.0, 3.0], B)
    elif kind == 3:    # tight curvature: 1 - e_y K(s) close to zero or negative (outside the model's domain)
        kp[:, 3] = rng.choice([-0.5, -0.2, 0.2, 0.5, 1.0], B); st[:, 1] = rng.choice([-2.5, -1.0, 1.0, 2.5, 4.0], B)
    elif kind == 4:    # random weights
        weights = [0.0] + [float(w) for w in rng.choice([0.0, 1e-3, 1.0, 9.0, 100.0, 1e4], 7)]
    return st, kp, vd, up, weights


def disagreements(g, o):
    """Indices where status differs, and where both are Optimal but the first move differs by more than 1e-5."""
    sm = np.nonzero(g["status"] != o["status"])[0]
    both = (g["status"] == 0) & (o["status"] == 0)
    du = np.abs(g["u0"] - o["u0"]).max(axis=1)
    return sm, np.nonzero(both & (du > 1e-5))[0], du


def main():
    from oracle import oracle as O
    import emu as E
    from mkz_mpc_path_follower_b200 import workload as W
    frenet = len(sys.argv) > 3 and sys.argv[3] == "frenet"
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    t0 = time.time()
    tot = bad = rnd = 0
    hist = {}
    B = 48
    while time.time() - t0 < budget:
        N = int(rng.choice([3, 5, 8, 8, 12, 20, 20, 31]))
        b0 = int(rng.integers(0, 1 << 30))
        if frenet:
            kind = rnd % FRENET_KINDS
            st, kp, vd, up, weights = perturb_frenet(rng, kind, W.make_frenet_batch(B, N, b0=b0), N)
            cfg = O.default_cfg_frenet(N, weights=weights)
            k = E.kcfg_from_oracle(cfg)
            o = O.solve_batch_frenet(cfg, st, kp, vd, up, n_threads=8)
            w = E.solve_batch_frenet(k, st, kp, vd, up)
            t = E.solve_batch_frenet(k, st, kp, vd, up, tpp=True, slots=17)
        else:
            kind = rnd % XY_KINDS
            b = W.make_batch(B, N, b0=b0)
            st, ref, vd, up, weights = perturb_xy(rng, kind, b, N)
            cfg = O.default_cfg(N, weights=weights)
            k = E.kcfg_from_oracle(cfg)
            warm = None
            if rnd % 3 == 2:   # warm start from the solution of the neighbouring problem
                w0 = O.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], want_traj=True, n_threads=8)["traj"]
                warm = np.roll(w0, 1, axis=0)
            cp = lambda a: None if a is None else a.copy()
            o = O.solve_batch(cfg, st, ref, vd, up, warm=cp(warm), n_threads=8)
            w = E.solve_batch(k, st, ref, vd, up, warm=cp(warm))
            t = E.solve_batch_tpp(k, st, ref, vd, up, warm=cp(warm), slots=17)
        tag = "frenet " if frenet else ""
        for name, g in (("warp", w), ("thread", t)):
            sm, dm, du = disagreements(g, o)
            for i in sm:
                print("round %d %skind %d N %d %s: problem %d status %d vs oracle %d (iters %d / %d)" % (
                    rnd, tag, kind, N, name, i, g["status"][i], o["status"][i], g["iters"][i], o["iters"][i]), flush=True)
            for i in dm:
                print("round %d %skind %d N %d %s: problem %d |du| %.3e (iters %d / %d, cost %.9g / %.9g)" % (
                    rnd, tag, kind, N, name, i, du[i], g["iters"][i], o["iters"][i], g["cost"][i], o["cost"][i]), flush=True)
            bad += sm.size + dm.size
        for s_ in o["status"]:
            hist[int(s_)] = hist.get(int(s_), 0) + 1
        tot += B
        rnd += 1
    print("%d problems in %d rounds, %d disagreements, oracle statuses %s, %.0f s" % (tot, rnd, bad, dict(sorted(hist.items())), time.time() - t0))


def main_rollout():
    """python tests/fuzz_layouts.py [seconds] [seed] rollout: closed loops of 12 control steps (fused rollout kernel's source and
    the per-period pipeline with the thread-per-problem solve, both emulated) against the oracle's closed loop, vehicle by vehicle:
    random poses around the three paths (one per round near the end: stop latch), time and distance mode, target speeds <= 0 too."""
    from oracle import oracle as O
    import emu as E
    from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    t0 = time.time()
    tot = bad = 0
    B, T = 6, 12
    while time.time() - t0 < budget:
        N = int(rng.choice([8, 8, 12, 20])); pid = int(rng.choice([1, 2, 3]))
        g = GPSRefTrajectory(mat_filename=pid, traj_horizon=N, traj_dt=0.2)
        tr = g.trajectory
        n = tr.shape[0]
        cfg = O.default_cfg(N)
        k = E.kcfg_from_oracle(cfg)
        seed = O.module_load_solution(cfg)
        j = rng.integers(0, n - 200, B); j[0] = n - rng.integers(2, 120)
        sc = rng.choice([0.1, 0.5, 2.0])
        poses = np.stack([tr[j, 4] + rng.normal(0, sc, B), tr[j, 5] + rng.normal(0, sc, B), tr[j, 3] + rng.normal(0, 0.1 * sc, B)], 1)
        mode = bool(rng.integers(0, 2)); vt = float(rng.choice([-1.0, 0.0, 3.0, 8.0]))
        path, keep = O.make_path(tr)
        for tpp in (False, True):
            log, final = E.rollout(k, tr, poses, T, track_using_time=mode, target_vel=vt, warm0=seed, tpp=tpp)
            for b in range(B):
                ol = O.closed_loop(cfg, path, poses[b], T, track_using_time=mode, target_vel=vt)
                m = len(ol)
                st_eq = np.array_equal(log[:m, b, 6], ol[:, 6])
                du = np.abs(log[:m, b, 4:6] - ol[:, 4:6]).max() if m else 0.0
                dp = np.abs(log[:m, b, 0:4] - ol[:, 0:4]).max() if m else 0.0
                tot += 1
                if not st_eq or du > 1e-5 or dp > 1e-6:
                    bad += 1
                    print("N %d path %d %s vehicle %d time mode %s target_vel %g spread %g: statuses equal %s, |du| %.2e, |dpose| %.2e, most iterations %d / %d" % (
                        N, pid, "pipeline" if tpp else "fused", b, mode, vt, sc, st_eq, du, dp, log[:m, b, 7].max(), ol[:, 7].max()), flush=True)
    print("%d closed loops, %d disagreements, %.0f s" % (tot, bad, time.time() - t0))


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "rollout":
        main_rollout()
    else:
        main()
