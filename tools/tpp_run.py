#!/usr/bin/env python
"""Thread-per-problem path only: `python tools/tpp_run.py [B] [N] [reps]` -> kernel ms per launch (for ncu and sweeps)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MPCB200_TPP_MIN_BATCH", "1")
from mkz_mpc_path_follower_b200 import capi, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
b = workload.make_batch(B, N)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("state", "ref", "u_prev", "v_des")}
s = capi.Solver(N)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
s.set_stream(st.cuda_stream)
u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev)
iters = torch.empty(B, dtype=torch.int32, device=dev)
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.solve_batch_device(B, d["state"], d["ref"], d["u_prev"], u0, v_des=d["v_des"], status=status, iters=iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = iters.cpu().numpy(); stt = status.cpu().numpy()
    print("rep %d: %.3f ms  B=%d N=%d  conv=%.5f  mean_iters=%.2f  solves/s=%.0f  [TPP_BLOCKS_PER_SM=%s]" %
          (r, ms, B, N, (stt == 0).mean(), it.mean(), (stt == 0).sum() / ms * 1e3, os.environ.get("MPCB200_TPP_BLOCKS_PER_SM", "max")), flush=True)
