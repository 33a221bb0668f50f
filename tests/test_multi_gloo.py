"""world_size-2 gloo test (CPU) of the multi-GPU host logic: row-shard ranges, global ids, the one
all-gather exchange and the merge contract, plus the replica-mode query partition."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exact(X, Q, k):
    d = ((Q[:, None, :].astype(np.float64) - X[None, :, :].astype(np.float64)) ** 2).sum(-1)
    idx = np.argsort(d, 1, kind="stable")[:, :k]
    return idx, np.take_along_axis(d, idx, 1)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from vectordb_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n, d, B, k = 1001, 8, 16, 5
    X = rng.random((n, d), dtype=np.float32)
    Q = rng.random((B, d), dtype=np.float32)
    lo, hi = sharded.shard_range(n, rank, world)
    li, ld = _exact(X[lo:hi], Q, k)                       # this rank's local top-k (local ids)
    if rank == 1:                                         # a ragged shard result: fewer than k valid entries
        li[0, 3:] = -1
        ld[0, 3:] = np.inf

    def merge_fn(all_i, all_d, kk):
        return sharded.numpy_merge(all_i.numpy(), all_d.numpy(), kk)

    gi, gd = sharded.exchange_and_merge(torch.from_numpy(li), torch.from_numpy(ld), lo, k, dist, merge_fn)
    if rank == 0:
        np.save(out, np.concatenate([gi.astype(np.float64), gd], 1))
    # replica mode: query partition covers every query exactly once
    ranges = [sharded.query_range(B, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == B and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_shard_exchange_gloo(tmp_path):
    world = 2
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    rng = np.random.default_rng(0)
    n, d, B, k = 1001, 8, 16, 5
    X = rng.random((n, d), dtype=np.float32)
    Q = rng.random((B, d), dtype=np.float32)
    ti, td = _exact(X, Q, k)
    gi, gd = got[:, :k].astype(np.int64), got[:, k:]
    # rows 1.. are exact; row 0 lost rank 1's entries beyond its 3rd (ragged shard) but must still be sorted & valid
    assert np.array_equal(gi[1:], ti[1:])
    assert np.allclose(gd[1:], td[1:])
    assert np.all(np.diff(gd, axis=1) >= 0)
    assert np.all(gi[0] >= 0)


def test_shard_ranges_and_global_ids():
    sys.path.insert(0, ROOT)
    from vectordb_b200 import sharded
    for n, w in ((10, 3), (100_000_000, 8), (7, 8)):
        rs = [sharded.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    ids = np.array([[0, 5, -1]])
    assert np.array_equal(sharded.to_global_ids(ids, 100), np.array([[100, 105, -1]]))
    i = np.array([[[3, 1, -1]], [[2, 7, 9]]])
    d = np.array([[[0.5, 0.7, np.inf]], [[0.5, 0.6, 0.9]]])
    mi, md = sharded.numpy_merge(i, d, 3)
    assert mi.tolist() == [[2, 3, 7]] and md.tolist() == [[0.5, 0.5, 0.6]]
