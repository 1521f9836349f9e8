#!/usr/bin/env python
"""bench.py -- converged MPC solves/s at batch 64K, N=20 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 5 --warmup 3
    python bench.py --impl reference ...      # CPU arm: the oracle (restated interior point, NOT Ipopt)

A step = one pass of the hot path (cold-start solve of every problem of the batch) over one
synthetic batch: BASELINE.json configs[2], 65,536 problems per GPU at N=20 along paths 1-3
(SURVEY.md 8d).  Weak scaling: every rank solves its own contiguous 65,536-problem slice of the
counter-based stream; the only inter-GPU traffic is one NCCL all-gather of the 32 B/problem
result record.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_RIC, F_EVAL = 1235, 310  # algorithmic flops per stage per interior-point iteration


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU")
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--x0", dest="start", default="zero", choices=["zero", "rollout"],
                    help="start point: zero = the reference's start=0.0 (the metric); rollout = MPCB200_START_ROLLOUT "
                         "(opt-in; what the horizon sweep of configs[4] uses at N = 40 / 80)")
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms 100` child for the whole region, parsed when it is stopped."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        n = 0
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            n += 1
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": n}


class NvmlClockSampler(object):
    """The same through NVML (pynvml), polled every 5 ms from a thread: the timed region is a few hundred ms, too short
    for more than a sample or two of `nvidia-smi -lms 100`.  Falls back to ClockSampler when NVML is not usable."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.fallback = None
        self.thread = None
        self.samples = []
        self.mask = 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            self.fallback = ClockSampler(index)

    def _reasons(self):
        for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            f = getattr(self.nv, fn, None)
            if f is not None:
                try:
                    return int(f(self.h))
                except Exception:
                    pass
        return 0

    def _run(self):
        while not self._stop:
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= self._reasons()
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.fallback is not None:
            return self.fallback.start()
        import threading
        self._stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.fallback is not None:
            return self.fallback.stop()
        self._stop = True
        self.thread.join(timeout=2)
        reasons = sorted(name for bit, name in self.BITS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml, 5 ms poll"}


def cpu_baseline(N, sample, threads, start_b0=0, start="zero"):
    """The oracle (restated CPU interior point, NOT Ipopt) on a bounded sample of the same workload."""
    from oracle import oracle as O
    from mkz_mpc_path_follower_b200 import workload
    O.build()
    b = workload.make_batch(sample, N, b0=start_b0)
    cfg = O.default_cfg(N)
    warm = O.rollout_start(cfg, b["state"], b["u_prev"]) if start == "rollout" else None
    t0 = time.perf_counter()
    r = O.solve_batch(cfg, b["state"], b["ref"], b["v_des"], b["u_prev"], warm=warm, n_threads=threads)
    dt = time.perf_counter() - t0
    conv = int((r["status"] == 0).sum())
    return conv / dt, dt, conv, r


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia/JuMP/Ipopt cannot run here, so
    this arm times the oracle port with every host core, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.horizon
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(256, 64 * cores)
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(N, min(sample, 256), cores, start=args.start)
    t_tot, conv_tot = 0.0, 0
    for s in range(args.steps):
        v, dt, conv, _ = cpu_baseline(N, sample, cores, start_b0=s * sample, start=args.start)
        t_tot += dt
        conv_tot += conv
    val = conv_tot / t_tot
    line = {
        "impl": "reference", "metric": "converged MPC solves/sec at batch 64K, N=20", "value": val,
        "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: N=%d cold-start solves along path1-3, bounded sample of %d problems/step" % (N, sample)},
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": "%d problems/step x %d steps, restated CPU interior point (oracle), NOT Ipopt" % (sample, args.steps)},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mkz_mpc_path_follower_b200 import capi, sharding, workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solver has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, B = args.horizon, args.batch
    nt = 6 * N + 4

    # ---- synthetic batch: this rank's contiguous slice of the problem stream
    b = workload.make_batch(B, N, b0=rank * B)
    solver = capi.Solver(N, device=local, start_mode=capi.START_ROLLOUT if args.start == "rollout" else capi.START_ZERO)
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream shared by torch and the library
    torch.cuda.set_stream(stream)
    solver.set_stream(stream.cuda_stream)
    d_state = torch.from_numpy(b["state"]).to(dev)
    d_ref = torch.from_numpy(b["ref"]).to(dev)
    d_uprev = torch.from_numpy(b["u_prev"]).to(dev)
    d_vdes = torch.from_numpy(b["v_des"]).to(dev)
    d_u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
    d_cost = torch.empty(B, dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    d_iters = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        flush.zero_()  # L2 flush between timed iterations (inputs are 36 MB < L2)
        solver.solve_batch_device(B, d_state, d_ref, d_uprev, d_u0, v_des=d_vdes, cost=d_cost, status=d_status, iters=d_iters)
        if world > 1:   # the one exchange step: all-gather of the 32 B/problem result record
            sharding.all_gather_records(sharding.pack_records(d_u0, d_cost, d_status, d_iters))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    # kernel-only time of one step (events directly around the kernel) for the roofline
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    sampler = NvmlClockSampler(local)
    sampler.start()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        flush.zero_()
        e0.record()
        solver.solve_batch_device(B, d_state, d_ref, d_uprev, d_u0, v_des=d_vdes, cost=d_cost, status=d_status, iters=d_iters)
        e1.record()
        if world > 1:   # the one exchange step: all-gather of the 32 B/problem result record
            sharding.all_gather_records(sharding.pack_records(d_u0, d_cost, d_status, d_iters))
        e1.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
    t1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = t0.elapsed_time(t1)
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    km = torch.zeros(world, dtype=torch.float64, device=dev)
    km[rank] = float(np.mean(kern_ms))
    if world > 1:
        dist.all_reduce(km, op=dist.ReduceOp.SUM)
    kernel_ms_per_rank = [round(float(v), 3) for v in km.tolist()]

    status = d_status.cpu().numpy(); iters = d_iters.cpu().numpy()
    conv_local = int((status == 0).sum())
    cnt = torch.tensor([conv_local, int(iters.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    conv_total, iters_total = float(cnt[0].item()), float(cnt[1].item())
    value = conv_total * args.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers (H2D + kernel + D2H timed)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    h_state, h_ref, h_uprev, h_vdes = pin(b["state"]), pin(b["ref"]), pin(b["u_prev"]), pin(b["v_des"])
    solver.set_stream(None)  # back to the handle's own stream
    for _ in range(2):
        solver.solve_batch(h_state, h_ref, h_uprev, v_des=h_vdes)
    barrier()
    w0 = time.perf_counter()
    h2d = d2h = launches = 0
    for _ in range(args.steps):
        r = solver.solve_batch(h_state, h_ref, h_uprev, v_des=h_vdes)
        st = solver.stats()
        h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
        launches += st["kernel_launches"]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = conv_total * args.steps / float(te.item())

    # ---- the same with the waypoints generated on the device (mpcb200_solve_batch_on_path): the host sends the
    # state, the path id and the previous command only (60 B/problem instead of 560)
    e2e_on_path = None
    if N <= 31:
        from mkz_mpc_path_follower_b200.gps_ref_traj import GPSRefTrajectory
        for i, pth in enumerate((1, 2, 3)):
            solver.set_path(i, GPSRefTrajectory(mat_filename=pth, traj_horizon=N, traj_dt=0.2).trajectory)
        h_pathof = torch.from_numpy((b["path"] - 1).astype(np.int32)).pin_memory().numpy()
        for _ in range(2):
            solver.solve_batch_on_path(h_state, h_pathof, h_uprev, v_des=h_vdes)
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            r2 = solver.solve_batch_on_path(h_state, h_pathof, h_uprev, v_des=h_vdes)
            st2 = solver.stats()
        torch.cuda.synchronize()
        tp = torch.tensor([time.perf_counter() - w0, float((r2["status"] == 0).sum())], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = tp[0:1].clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            csum = tp[1:2].clone(); dist.all_reduce(csum, op=dist.ReduceOp.SUM)
            tp = torch.cat((tmax, csum))
        e2e_on_path = {"value": float(tp[1].item()) * args.steps / float(tp[0].item()), "unit": "solves/s",
                       "h2d_bytes_per_step": int(st2["h2d_bytes"]), "d2h_bytes_per_step": int(st2["d2h_bytes"]),
                       "what": "mpcb200_solve_batch_on_path: reference waypoints generated on the device"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel: FP64 CUDA-core pipe, SURVEY 8(d)
    fp64_peak = solver.fp64_peak_tflops()
    k_ms = float(np.mean(kern_ms))
    flops = float(iters.sum()) * (F_RIC + F_EVAL) * N
    achieved = flops / (k_ms * 1e-3) / 1e12
    io_bytes = B * ((4 + 2 + 3 * (N + 1) + 1) * 8 + 3 * 8 + 2 * 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    tr_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tr_file) and N == 20 and B == 65536:   # the capture is of this configuration
        try:
            traffic = json.load(open(tr_file)).get("bytes_per_launch")
        except Exception:
            pass
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
        "traffic": traffic,
        "peak_source": "FP64 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
        "flops_model": "sum_p iters_p * (1235 + 310) * N, SURVEY.md 8(d)",
        "kernel": "mpc_solve_kernel", "kernel_ms": k_ms, "kernel_ms_per_rank": kernel_ms_per_rank,
        "hbm": {"achieved": io_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": io_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback of B200_PROFILING.md",
                "algorithmic_bytes_per_solve": io_bytes // B},
    }

    # ---- CPU baseline beside it: oracle on a bounded sample, all cores
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(256, 48 * cores)
    cpu_v, cpu_dt, cpu_conv, cpu_r = cpu_baseline(N, sample, cores, start=args.start)
    ok = (cpu_r["status"] == 0) & (status[:sample] == 0) if rank == 0 else None
    u0 = d_u0.cpu().numpy()
    parity = {
        "sample": sample, "status_equal": bool((cpu_r["status"] == status[:sample]).all()),
        "max_abs_du": float(np.abs(u0[:sample] - cpu_r["u0"])[ok].max()) if ok.any() else None,
    }

    line = {
        "metric": "converged MPC solves/sec at batch 64K, N=20" if (N == 20 and B == 65536) else "converged MPC solves/sec at batch %d per GPU, N=%d" % (B, N),
        "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[%d]: %d solves per GPU from the %s, N=%d, paths 1-3 round-robin, "
                               "perturbed states (SURVEY 8d)" % (2 if (N == 20 and args.start == "zero") else 4, B,
                                                                 "all-zero start (start=0.0)" if args.start == "zero" else "rollout start (opt-in)", N),
                   "batch_per_gpu": B, "horizon": N, "l2": "256 MiB flush between steps (inputs 36 MB < L2)",
                   "start": args.start, "max_iter": int(solver.cfg.max_iter)},
        "converged_frac": conv_total / (world * B), "mean_iters": iters_total / (world * B),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_v, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": "first %d problems of the same batch, restated CPU interior point (oracle), NOT Ipopt" % sample},
        "parity_vs_oracle": parity,
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "e2e_on_path": e2e_on_path,
        "gpu_launches": int(args.steps * 1),
        "clocks": clocks,
    }

    if N <= 31:
        # the Frenet-frame variant (MKZMPCPathFollowerFrenet.jl) on the same batch size and horizon, outside the
        # timed region of the headline metric: device-resident inputs, CUDA events on the shared stream, oracle on a sample
        try:
            fb = workload.make_frenet_batch(B, N, b0=rank * B)
            fs = capi.FrenetSolver(N, device=local)
            fs.set_stream(stream.cuda_stream)
            fd = {k_: torch.from_numpy(fb[k_]).to(dev) for k_ in ("state", "kpoly", "u_prev", "v_des")}
            f_u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); f_st = torch.empty(B, dtype=torch.int32, device=dev)
            f_it = torch.empty(B, dtype=torch.int32, device=dev)
            f_ms = []
            for i in range(3 + args.steps):
                flush.fill_(1)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fs.solve_batch_device(B, fd["state"], fd["kpoly"], fd["u_prev"], f_u0, v_des=fd["v_des"], status=f_st, iters=f_it)
                e1.record(stream); torch.cuda.synchronize()
                if i >= 3:
                    f_ms.append(e0.elapsed_time(e1))
            fst = f_st.cpu().numpy(); fit = f_it.cpu().numpy()
            from oracle import oracle as O
            n_s = min(B, 512)
            fo = O.solve_batch_frenet(O.default_cfg_frenet(N, max_iter=int(fs.cfg.max_iter)), fb["state"][:n_s], fb["kpoly"][:n_s],
                                      fb["v_des"][:n_s], fb["u_prev"][:n_s], n_threads=cores)
            fok = (fo["status"] == 0) & (fst[:n_s] == 0)
            line["frenet_variant"] = {
                "value": float((fst == 0).sum()) / (float(np.mean(f_ms)) * 1e-3), "unit": "solves/s", "kernel": "mpc_solve_frenet_kernel",
                "kernel_ms": float(np.mean(f_ms)), "converged_frac": float((fst == 0).mean()), "mean_iters": float(fit.mean()),
                "fp64_tflops": float(fit.astype(np.float64).sum()) * (F_RIC + F_EVAL) * N / (float(np.mean(f_ms)) * 1e-3) / 1e12,
                "parity_vs_oracle": {"sample": n_s, "status_equal": bool((fo["status"] == fst[:n_s]).all()),
                                     "max_abs_du": float(np.abs(f_u0.cpu().numpy()[:n_s] - fo["u0"])[fok].max())},
                "what": "same batch size and horizon, rank 0's slice, all-zero start, workload.make_frenet_batch"}
        except Exception as ex:   # the headline line must not depend on the variant
            line["frenet_variant"] = {"error": repr(ex)}

    if not args.no_latency:
        # single-solve latency through the C ABI (host buffers, H2D + kernel + D2H), warm-started like the
        # control loop; CPU oracle beside it
        from oracle import oracle as O
        s1 = capi.Solver(N, device=local)
        ocfg = O.default_cfg(N)
        nl = 264
        sol = s1.solve_batch(b["state"][:64], b["ref"][:64], b["u_prev"][:64], v_des=b["v_des"][:64], want_traj=True)["traj"]
        lat = {"g_warm": [], "c_warm": [], "g_cold": [], "c_cold": []}
        for i in range(nl):
            j = i % 64
            for mode in ("warm", "cold"):
                # warm: from the problem's own previous solution, like the control loop re-solving 0.1 s later; cold: start=0.0
                wg = sol[j:j + 1].copy() if mode == "warm" else None
                wc = sol[j].copy() if mode == "warm" else None
                a = time.perf_counter()
                s1.solve_batch(b["state"][j:j + 1], b["ref"][j:j + 1], b["u_prev"][j:j + 1], v_des=b["v_des"][j:j + 1], warm=wg)
                lat["g_" + mode].append(time.perf_counter() - a)
                if i < 96:   # the CPU leg is slow from the cold start: a bounded sample
                    a = time.perf_counter()
                    O.solve(ocfg, b["state"][j], b["ref"][j], 1.0, b["u_prev"][j], warm=wc)
                    lat["c_" + mode].append(time.perf_counter() - a)
        pct = lambda v, q: 1e3 * float(np.percentile(v[8:], q))
        line["latency"] = {"gpu_p50_ms": pct(lat["g_warm"], 50), "gpu_p99_ms": pct(lat["g_warm"], 99),
                           "cpu_oracle_p50_ms": pct(lat["c_warm"], 50),
                           "gpu_cold_p50_ms": pct(lat["g_cold"], 50), "gpu_cold_p99_ms": pct(lat["g_cold"], 99),
                           "cpu_oracle_cold_p50_ms": pct(lat["c_cold"], 50),
                           "what": "batch of 1 through the C ABI incl. H2D/D2H (one mpcb200_solve_batch call per solve); warm = start from "
                                   "the problem's previous solution like the control loop, cold = the all-zero start; the oracle is the "
                                   "restated CPU interior point on one thread, NOT Ipopt; closed-loop percentiles: tools/closed_loop_latency.py"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
