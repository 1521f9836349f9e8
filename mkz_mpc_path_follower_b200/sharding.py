"""Multi-GPU plumbing for the batched solver.

Problems are independent, so the batch is cut into contiguous slices, one per rank (one process
per GPU); nothing crosses GPUs during the solve.  The single exchange step is one all-gather of the
per-problem result record {acc, df, cost: 3 x f64; status, iters: 2 x i32} = 32 B, done with
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np

RECORD_DOUBLES = 4  # 32 bytes


def shard_range(total, world, rank):
    """Contiguous slice [lo, hi) of `total` problems owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(u0, cost, status, iters):
    """torch tensors (any device): u0 (B,2) f64, cost (B,) f64, status/iters (B,) i32 -> (B,4) f64
    whose last column carries the two int32 bit patterns."""
    import torch
    B = u0.shape[0]
    rec = torch.empty((B, RECORD_DOUBLES), dtype=torch.float64, device=u0.device)
    rec[:, 0:2] = u0
    rec[:, 2] = cost
    rec[:, 3] = torch.stack((status, iters), dim=1).contiguous().view(torch.float64).squeeze(1)
    return rec


def unpack_records(rec):
    """Inverse of pack_records on a (B,4) f64 torch tensor -> u0, cost, status, iters."""
    import torch
    si = rec[:, 3].contiguous().view(torch.int32).view(-1, 2)
    return rec[:, 0:2], rec[:, 2], si[:, 0], si[:, 1]


def unpack_records_np(rec):
    """The same for a (B,4) float64 numpy array of records written by the solve kernel (mpcb200_solve_batch_records):
    returns u0, cost, status, iters, restorations (status word: low 8 bits status, bits 8..15 restoration count)."""
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    si = rec[:, 3].copy().view(np.int32).reshape(-1, 2)
    return rec[:, 0:2], rec[:, 2], si[:, 0] & 0xFF, si[:, 1], (si[:, 0] >> 8) & 0xFF


def all_gather_records(rec, sizes=None):
    """One all-gather of the result records.  Equal slices use all_gather_into_tensor; ragged slices
    (total not divisible by world) are padded to the largest slice and trimmed."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    if sizes is None:
        sizes = [rec.shape[0]] * world
    mx = max(sizes)
    if rec.shape[0] < mx:
        pad = torch.zeros((mx - rec.shape[0], RECORD_DOUBLES), dtype=rec.dtype, device=rec.device)
        rec = torch.cat((rec, pad), dim=0)
    out = torch.empty((world * mx, RECORD_DOUBLES), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)
