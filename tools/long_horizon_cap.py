#!/usr/bin/env python
"""configs[4] from the reference's all-zero start (`start=0.0`) under Ipopt's own iteration cap (max_iter = 3000, the 3.12
default) instead of the library's 200: converged fraction, status counts and the iteration histogram, per horizon, for both
device layouts of the solver, and the oracle on a sample.
    python tools/long_horizon_cap.py [B40] [B80] [oracle_sample]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)

B40 = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B80 = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
n_or = int(sys.argv[3]) if len(sys.argv) > 3 else 128
CAP = 3000
dev = torch.device("cuda", 0)
out = {}
for N, B in ((20, B40), (40, B40), (80, B80)):
    b = workload.make_batch(B, N)
    d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
    row = {"problems": B, "max_iter": CAP}
    res = {}
    for name, mb in (("warp_per_problem", 0), ("thread_per_problem", 1)):
        s = capi.Solver(config=capi.default_config(N, max_iter=CAP))
        s.set_large_batch_path(mb)
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        s.set_stream(stream.cuda_stream)
        u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], status=status, iters=iters)
        e1.record()
        torch.cuda.synchronize()
        st = status.cpu().numpy(); it = iters.cpu().numpy()
        res[name] = (u0.cpu().numpy(), st, it)
        conv = st == 0
        row[name] = {"kernel_ms": e0.elapsed_time(e1), "converged_frac": float(conv.mean()),
                     "status_counts": {str(k): int((st == k).sum()) for k in range(5)},
                     "iters_of_converged": {"p50": float(np.percentile(it[conv], 50)), "p90": float(np.percentile(it[conv], 90)),
                                            "p99": float(np.percentile(it[conv], 99)), "max": int(it[conv].max())} if conv.any() else None,
                     "converged_within": {str(c): float((conv & (it <= c)).mean()) for c in (100, 200, 400, 800, 1600, 3000)},
                     "restorations": int(np.asarray(s.restorations(B)).sum()) if mb == 0 else None}
        s.close()
    (uw, sw, iw), (ut, st_, it_) = res["warp_per_problem"], res["thread_per_problem"]
    both = (sw == 0) & (st_ == 0)
    row["layouts_agree"] = {"status_equal": float((sw == st_).mean()), "both_converged": float(both.mean()),
                            "max_abs_du_both_converged": float(np.abs(uw - ut)[both].max()) if both.any() else None,
                            "frac_du_gt_1e-5": float((np.abs(uw - ut)[both].max(axis=1) > 1e-5).mean()) if both.any() else None}
    if n_or:
        t0 = time.time()
        o = O.solve_batch(O.default_cfg(N, max_iter=CAP), b["state"][:n_or], b["ref"][:n_or], b["v_des"][:n_or], b["u_prev"][:n_or], n_threads=16)
        bo = (o["status"] == 0) & (sw[:n_or] == 0)
        row["oracle_sample"] = {"problems": n_or, "seconds": time.time() - t0, "oracle_converged_frac": float((o["status"] == 0).mean()),
                                "gpu_converged_frac_same_problems": float((sw[:n_or] == 0).mean()), "status_equal": float((o["status"] == sw[:n_or]).mean()),
                                "max_abs_du_both_converged": float(np.abs(o["u0"] - uw[:n_or])[bo].max()) if bo.any() else None,
                                "frac_du_gt_1e-5": float((np.abs(o["u0"] - uw[:n_or])[bo].max(axis=1) > 1e-5).mean()) if bo.any() else None,
                                "iters_equal": float((o["iters"] == iw[:n_or]).mean())}
    out["N%d" % N] = row
    print(json.dumps({"N%d" % N: row}), flush=True)
