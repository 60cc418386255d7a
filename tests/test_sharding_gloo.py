"""World-size-2 gloo test (CPU) of the multi-GPU sharding algebra: contiguous keypoint shards + one all-reduce (sum)
of the packed accumulator reproduce the unsharded normal equations. The per-shard sums come from the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ct_icp_b200 import _abi as abi
from ct_icp_b200.sharding import pack_normal_equations, shard_bounds, unpack_normal_equations


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from test_golden import _load_case
    return _load_case()


def _raw_sums(orc, z, kp, frame):
    """Un-normalised Σ u u^T, Σ -u s and count of a keypoint shard (no motion model → no regulariser)."""
    m = orc.voxel_map(orc.legacy_map_options(1.0, 20, 0.1))
    m.insert(z["map_xyz"])
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    io.min_number_neighbors = 10
    A, b, n = m.gn_normal_equations(io, np.ascontiguousarray(kp), frame)
    if n < 100:       # the oracle normalises only when n >= 100 (ct_icp.cpp:860-882)
        return A, b, n
    return A * n, b * n, n


def _worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from oracle_lib import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle()
    z, kp, frame, _ = _case()
    lo, hi = shard_bounds(len(kp), rank, world)
    A, b, n = _raw_sums(orc, z, kp[lo:hi], frame)
    acc = torch.from_numpy(pack_normal_equations(A, b, n))
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_partition():
    for K in (0, 1, 7, 1215, 100003):
        for G in (1, 2, 3, 8):
            spans = [shard_bounds(K, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == K
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_allreduce_of_shards_equals_unsharded(orc, tmp_path):
    z, kp, frame, _ = _case()
    A_full, b_full, n_full = _raw_sums(orc, z, kp, frame)
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    A, b, n = unpack_normal_equations(np.load(out))
    assert n == n_full
    assert np.abs(A - A_full).max() < 1e-9 * np.abs(A_full).max()
    assert np.abs(b - b_full).max() < 1e-9 * max(np.abs(b_full).max(), 1e-6)


def _prefix_worker(rank, world, port, out, limit):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from ct_icp_b200.sharding import prefix_shares
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    valid = np.random.default_rng(11).random(1001) < 0.6            # the same flags on every rank
    lo, hi = shard_bounds(len(valid), rank, world)
    one_hot = torch.zeros(world, dtype=torch.float64)
    one_hot[rank] = float(valid[lo:hi].sum())
    dist.all_reduce(one_hot, op=dist.ReduceOp.SUM)                   # = all-gather of the per-rank counts
    before, share, total = prefix_shares(one_hot.numpy(), rank, limit)
    mine = (lo + np.flatnonzero(valid[lo:hi]))[:share]               # this rank's first `share` valid keypoints
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.tolist(), total))
    if rank == 0:
        np.save(out, np.array(sorted(sum((g[0] for g in gathered), [])) + [gathered[0][1]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("limit", [-1, 300, 450, 5000])
def test_sharded_residual_prefix_is_the_global_prefix(tmp_path, limit):
    """max_num_residuals is a PREFIX rule (ct_icp.cpp:409-424): the union of the per-rank shares must be exactly the
    first `limit` valid keypoints, whichever rank they live on (k_lm_select mode 0/1)."""
    out = str(tmp_path / "sel.npy")
    mp.spawn(_prefix_worker, args=(2, _free_port(), out, limit), nprocs=2, join=True)
    got = np.load(out)
    valid = np.random.default_rng(11).random(1001) < 0.6
    want = np.flatnonzero(valid)[: (limit if limit > 0 else None)]
    assert got[-1] == len(want)
    assert np.array_equal(got[:-1], want)
