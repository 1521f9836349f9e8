"""Throughput of the Frenet-frame variant (mpcb200_solve_batch_frenet) on one GPU: B cold solves at horizon N from
the all-zero start, device-resident inputs, CUDA-event timing on the handle's stream; the oracle on the host cores
beside it.  python tools/frenet_bench.py [--batch 65536] [--horizon 20] [--steps 5]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536); ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    from mkz_mpc_path_follower_b200 import capi, workload
    B, N = a.batch, a.horizon
    b = workload.make_frenet_batch(B, N)
    dev = torch.device("cuda:0")
    s = capi.FrenetSolver(N)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); s.set_stream(stream.cuda_stream)
    d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "kpoly", "u_prev", "v_des")}
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); cost = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev); iters = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def step():
        s.solve_batch_device(B, d["state"], d["kpoly"], d["u_prev"], u0, v_des=d["v_des"], cost=cost, status=status, iters=iters)
    for _ in range(max(3, a.warmup)):
        step()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.steps):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); step(); e1.record(stream); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    st = status.cpu().numpy(); it = iters.cpu().numpy()
    conv = int((st == 0).sum()); k_ms = float(np.mean(ms))
    # end to end through the host API
    t0 = time.perf_counter()
    for _ in range(a.steps):
        g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"])
    e2e = a.steps * int((g["status"] == 0).sum()) / (time.perf_counter() - t0)
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n_cpu = min(B, 64 * cores)
    t0 = time.perf_counter()
    o = O.solve_batch_frenet(O.default_cfg_frenet(N, max_iter=s.cfg.max_iter), b["state"][:n_cpu], b["kpoly"][:n_cpu], b["v_des"][:n_cpu],
                             b["u_prev"][:n_cpu], n_threads=cores)
    cpu = int((o["status"] == 0).sum()) / (time.perf_counter() - t0)
    ok = (o["status"] == 0) & (st[:n_cpu] == 0)
    flops = float(it.astype(np.float64).sum()) * (1235 + 310) * N
    print(json.dumps({"metric": "converged Frenet-variant MPC solves/sec", "value": conv / (k_ms * 1e-3), "unit": "solves/s", "batch": B, "horizon": N,
                      "kernel_ms": k_ms, "converged_frac": conv / B, "mean_iters": float(it.mean()), "e2e": e2e,
                      "fp64_tflops_model": flops / (k_ms * 1e-3) / 1e12,
                      "cpu_baseline": {"value": cpu, "cores": cores, "kind": "port", "sample": n_cpu},
                      "parity_vs_oracle": {"sample": n_cpu, "status_equal": bool((o["status"] == st[:n_cpu]).all()),
                                           "max_abs_du": float(np.abs(u0.cpu().numpy()[:n_cpu] - o["u0"])[ok].max())}}))


if __name__ == "__main__":
    main()
