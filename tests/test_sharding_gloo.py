"""N > 1 path on the CPU: two gloo ranks each take their contiguous slice of one synthetic batch,
"solve" it (with the oracle standing in for the GPU, which is absent here), all-gather the 32-byte
records and must reproduce the single-rank result bit for bit, in problem order."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, N, out_dir, frenet=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mkz_mpc_path_follower_b200 import sharding, workload
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(total, world, rank)
    if frenet:   # the Frenet-frame variant shards the same way (its workload is a counter-based stream too)
        b = workload.make_frenet_batch(hi - lo, N, b0=lo)
        r = O.solve_batch_frenet(O.default_cfg_frenet(N), b["state"], b["kpoly"], b["v_des"], b["u_prev"])
    else:
        b = workload.make_batch(hi - lo, N, b0=lo)
        r = O.solve_batch(O.default_cfg(N), b["state"], b["ref"], b["v_des"], b["u_prev"])
    rec = sharding.pack_records(torch.from_numpy(r["u0"]), torch.from_numpy(r["cost"]),
                                torch.from_numpy(r["status"]), torch.from_numpy(r["iters"]))
    sizes = [sharding.shard_range(total, world, q)[1] - sharding.shard_range(total, world, q)[0] for q in range(world)]
    allrec = sharding.all_gather_records(rec, sizes)
    u0, cost, status, iters = sharding.unpack_records(allrec)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), u0=u0.numpy(), cost=cost.numpy(), status=status.numpy(), iters=iters.numpy())
    dist.destroy_process_group()


def test_shard_ranges():
    from mkz_mpc_path_follower_b200 import sharding
    for total, world in ((65536, 8), (10, 3), (7, 8), (0, 2)):
        cuts = [sharding.shard_range(total, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        sz = [b - a for a, b in cuts]
        assert max(sz) - min(sz) <= 1


@pytest.mark.parametrize("total", [24, 25])   # even and ragged split
def test_two_rank_gather_equals_single_rank(tmp_path, total):
    import torch.multiprocessing as mp
    from mkz_mpc_path_follower_b200 import workload
    from oracle import oracle as O
    N, world = 8, 2
    port = 29500 + (os.getpid() % 2000) + total
    mp.spawn(_worker, args=(world, port, total, N, str(tmp_path)), nprocs=world, join=True)
    b = workload.make_batch(total, N)
    ref = O.solve_batch(O.default_cfg(N), b["state"], b["ref"], b["v_des"], b["u_prev"])
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(g["u0"], ref["u0"]) and np.array_equal(g["cost"], ref["cost"])
        assert np.array_equal(g["status"], ref["status"]) and np.array_equal(g["iters"], ref["iters"])


def test_two_rank_gather_equals_single_rank_frenet(tmp_path):
    import torch.multiprocessing as mp
    from mkz_mpc_path_follower_b200 import workload
    from oracle import oracle as O
    N, world, total = 8, 2, 21
    port = 29500 + (os.getpid() % 2000) + 77
    mp.spawn(_worker, args=(world, port, total, N, str(tmp_path), True), nprocs=world, join=True)
    b = workload.make_frenet_batch(total, N)
    ref = O.solve_batch_frenet(O.default_cfg_frenet(N), b["state"], b["kpoly"], b["v_des"], b["u_prev"])
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(g["u0"], ref["u0"]) and np.array_equal(g["cost"], ref["cost"])
        assert np.array_equal(g["status"], ref["status"]) and np.array_equal(g["iters"], ref["iters"])


def test_shard_range_properties():
    """Contiguous, disjoint, covering, balanced to within one problem -- for any total and world size."""
    from hypothesis import given, settings, strategies as st
    from mkz_mpc_path_follower_b200 import sharding

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 2_000_000), st.integers(1, 64))
    def check(total, world):
        r = [sharding.shard_range(total, world, q) for q in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(r[q][1] == r[q + 1][0] for q in range(world - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0

    check()


def test_kernel_record_layout_roundtrip():
    """The 32-byte record the solve kernel writes (status | restorations << 8 in the low int32 of the fourth double, iters in
    the high one) unpacks to the same fields pack_records / unpack_records move."""
    import torch
    from mkz_mpc_path_follower_b200 import sharding
    rng = np.random.default_rng(1)
    B = 257
    u0 = rng.normal(size=(B, 2)); cost = rng.normal(size=B) ** 2
    status = rng.integers(0, 5, B).astype(np.int32); iters = rng.integers(0, 3000, B).astype(np.int32)
    resto = rng.integers(0, 4, B).astype(np.int32)
    rec = sharding.pack_records(torch.from_numpy(u0), torch.from_numpy(cost), torch.from_numpy(status | (resto << 8)), torch.from_numpy(iters))
    a, c, s, i, r = sharding.unpack_records_np(rec.numpy())
    assert np.array_equal(a, u0) and np.array_equal(c, cost) and np.array_equal(s, status) and np.array_equal(i, iters) and np.array_equal(r, resto)
    # the C library's slice rule is the Python one
    import ctypes as C
    assert [sharding.shard_range(65537, 8, q) for q in range(8)][3] == (24577, 32769)
