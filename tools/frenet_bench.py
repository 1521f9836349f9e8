"""Throughput of the Frenet-frame variant (mpcb200_solve_batch_frenet): B cold solves per GPU at horizon N from the all-zero
start, device-resident inputs, CUDA-event timing on the handle's stream (max over ranks), the oracle on the host cores beside
it.  One GPU:  python tools/frenet_bench.py [--batch 65536] [--horizon 20] [--steps 5]
N GPUs (weak scaling, contiguous slices of the problem stream, one all-gather of the 32-byte result records per step):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/frenet_bench.py"""
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
    import torch.distributed as dist
    from mkz_mpc_path_follower_b200 import capi, workload, sharding
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = a.batch, a.horizon
    b = workload.make_frenet_batch(B, N, b0=rank * B)
    s = capi.FrenetSolver(N, device=local)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); s.set_stream(stream.cuda_stream)
    d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "kpoly", "u_prev", "v_des")}
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); cost = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev); iters = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        s.solve_batch_device(B, d["state"], d["kpoly"], d["u_prev"], u0, v_des=d["v_des"], cost=cost, status=status, iters=iters)
        if world > 1:
            return sharding.all_gather_records(sharding.pack_records(u0, cost, status, iters))
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        step()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(a.steps):
        flush.fill_(1)
        allrec = step()
    t1.record(stream)
    barrier()
    tm = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    conv = torch.tensor([float((status == 0).sum().item()), float(iters.sum().item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX); dist.all_reduce(conv, op=dist.ReduceOp.SUM)
    ms = float(tm.item()) / a.steps
    if rank == 0:
        if world > 1:   # the gathered records hold every rank's statuses, in problem order
            _, _, st_all, _ = sharding.unpack_records(allrec)
            assert int((st_all == 0).sum().item()) == int(conv[0].item())
        st = status.cpu().numpy()
        t0w = time.perf_counter()
        for _ in range(a.steps):
            g = s.solve_batch(b["state"], b["kpoly"], b["u_prev"], v_des=b["v_des"])
        e2e = a.steps * int((g["status"] == 0).sum()) / (time.perf_counter() - t0w)
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        n_cpu = min(B, 64 * cores)
        t0w = time.perf_counter()
        o = O.solve_batch_frenet(O.default_cfg_frenet(N, max_iter=s.cfg.max_iter), b["state"][:n_cpu], b["kpoly"][:n_cpu], b["v_des"][:n_cpu],
                                 b["u_prev"][:n_cpu], n_threads=cores)
        cpu = int((o["status"] == 0).sum()) / (time.perf_counter() - t0w)
        ok = (o["status"] == 0) & (st[:n_cpu] == 0)
        print(json.dumps({"metric": "converged Frenet-variant MPC solves/sec", "value": float(conv[0].item()) / (ms * 1e-3), "unit": "solves/s",
                          "n_gpus": world, "batch_per_gpu": B, "horizon": N, "ms_per_step": ms, "scaling": "weak",
                          "converged_frac": float(conv[0].item()) / (world * B), "mean_iters": float(conv[1].item()) / (world * B),
                          "e2e_rank0": e2e, "fp64_tflops_model": float(conv[1].item()) * (1235 + 310) * N / (ms * 1e-3) / 1e12,
                          "cpu_baseline": {"value": cpu, "cores": cores, "kind": "port", "sample": n_cpu},
                          "parity_vs_oracle": {"sample": n_cpu, "status_equal": bool((o["status"] == st[:n_cpu]).all()),
                                               "max_abs_du": float(np.abs(u0.cpu().numpy()[:n_cpu] - o["u0"])[ok].max())}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
