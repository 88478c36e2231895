"""CPU, gloo, world_size 2: the N > 1 host logic (phet_b200/distributed.py) -- balanced shards,
rank-0 draw + broadcast, and that per-rank partial accumulators all-reduce to the single-rank
result.  The oracle stands in for the device kernels here (tests may use it as the checker)."""
import os
import socket

import numpy as np
import pytest

from phet_b200.distributed import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 125, 1000, 30000):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import phet_oracle as po
    from phet_b200.distributed import allreduce_numpy, broadcast_numpy, shard_range
    rng = np.random.default_rng(0)
    N, G, S = 120, 40, 21
    y = np.zeros(N, dtype=np.int64)
    y[60:] = 1
    X = rng.gamma(2.0, 1.0, (N, G))
    m = po.subsample_size(y)
    # rank 0 draws, everyone receives the same index sets (what PHet.fit_predict does)
    if rank == 0:
        np.random.seed(12345)
        idx_i, idx_j, _ = po.draw_index_sets(y, m, S)
        idx = np.stack([idx_i, idx_j], axis=1).astype(np.int32)
    else:
        idx = np.zeros((S, 2, m), dtype=np.int32)
    idx = broadcast_numpy(idx, dist)
    Xz = po.zscore_nan_to_num(X)
    lo, hi = shard_range(S, rank, world)
    P, R = po.step1_all(Xz, idx[lo:hi, 0], idx[lo:hi, 1], (25, 75), m)
    acc = np.stack([po.fisher_terms(P).sum(1), R.sum(1)])
    acc = allreduce_numpy(acc, dist)
    # Step 3 shards genes; ranks contribute zeros outside their slice
    g_lo, g_hi = shard_range(G, rank, world)
    ctrl, case = np.where(y == 0)[0], np.where(y == 1)[0]
    prng = np.random.RandomState(5)
    perm_i = np.stack([prng.permutation(len(ctrl)) for _ in range(G)])
    perm_j = np.stack([prng.permutation(len(case)) for _ in range(G)])
    o = np.zeros(G)
    H, slices = po.step3_h(Xz[:, g_lo:g_hi], ctrl, case, perm_i[g_lo:g_hi], perm_j[g_lo:g_hi], m)
    o[g_lo:g_hi] = po.step3_o_from_h(H, slices)
    o = allreduce_numpy(o, dist)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), acc=acc, o=o, idx=idx)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle import phet_oracle as po
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["idx"], r1["idx"])
    assert np.array_equal(r0["acc"], r1["acc"]) and np.array_equal(r0["o"], r1["o"])
    rng = np.random.default_rng(0)
    N, G, S = 120, 40, 21
    y = np.zeros(N, dtype=np.int64)
    y[60:] = 1
    X = rng.gamma(2.0, 1.0, (N, G))
    Xz = po.zscore_nan_to_num(X)
    m = po.subsample_size(y)
    idx = r0["idx"]
    P, R = po.step1_all(Xz, idx[:, 0], idx[:, 1], (25, 75), m)
    assert np.allclose(r0["acc"][0], po.fisher_terms(P).sum(1), rtol=1e-13)
    assert np.allclose(r0["acc"][1], R.sum(1), rtol=1e-13)
    ctrl, case = np.where(y == 0)[0], np.where(y == 1)[0]
    prng = np.random.RandomState(5)
    perm_i = np.stack([prng.permutation(len(ctrl)) for _ in range(G)])
    perm_j = np.stack([prng.permutation(len(case)) for _ in range(G)])
    H, slices = po.step3_h(Xz, ctrl, case, perm_i, perm_j, m)
    assert np.array_equal(r0["o"], po.step3_o_from_h(H, slices))
