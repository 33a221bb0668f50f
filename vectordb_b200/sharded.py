"""Host logic of the multi-GPU deployment of the search path.

Two partitionings, one process per GPU (torch.distributed for the plumbing):
  * replicas (default when the table fits one GPU's HBM): every rank holds the whole table and its graph;
    the QUERY stream is partitioned — independent units, no data-path collective.
  * row shards (the table exceeds one GPU: config C5, 100M x 768 over 8 GPUs): rank r holds rows
    [r*n/S, (r+1)*n/S) with its own graph over LOCAL ids; the query batch is replicated; every rank returns
    its top-k with GLOBAL ids (base_r + local id); ONE all-gather of [B x k] (id, dist) per rank followed by a
    k-way merge by (distance, id) on every rank.  Exact: each shard must return its own top-k, not k/S.
    The single-segment reference has the same two-source merge between graph and tail results
    (engine/db/execution/vec_search_executor.cpp:885-900).
The merge itself is a CUDA kernel (eps_merge_shards_device); `merge_fn` is injectable so the plumbing can be
exercised on CPU with the gloo backend.
"""
import numpy as np


def shard_range(n_rows, rank, world):
    """Contiguous row range of `rank` (the last shard takes the remainder)."""
    per = n_rows // world
    lo = rank * per
    hi = n_rows if rank == world - 1 else lo + per
    return lo, hi


def query_range(n_queries, rank, world):
    """Replica mode: the slice of the query stream rank `rank` serves."""
    per = (n_queries + world - 1) // world
    lo = min(n_queries, rank * per)
    return lo, min(n_queries, lo + per)


def to_global_ids(local_ids, base):
    """local row ids -> global ids; -1 (empty slot) stays -1.  Works on numpy arrays and torch tensors."""
    return (local_ids + base) * (local_ids >= 0) + (-1) * (local_ids < 0)


def numpy_merge(ids, dists, k):
    """Checker / CPU stand-in for the merge kernel: ids, dists [S, B, k] -> [B, k] by (distance, id)."""
    S, B, kk = ids.shape
    cat_i = np.transpose(ids, (1, 0, 2)).reshape(B, S * kk)
    cat_d = np.transpose(dists, (1, 0, 2)).reshape(B, S * kk).astype(np.float64)
    cat_d = np.where(cat_i < 0, np.inf, cat_d)
    big = np.where(cat_i < 0, np.iinfo(np.int64).max, cat_i)
    order = np.lexsort((big, cat_d), axis=1)[:, :k]
    oi = np.take_along_axis(cat_i, order, 1)
    od = np.take_along_axis(cat_d, order, 1)
    return np.where(np.isinf(od), -1, oi), od


def exchange_and_merge(local_ids, local_dists, base, k, dist, merge_fn):
    """All-gather this rank's [B, k] (global id, dist) and merge.  `dist` is torch.distributed (any backend);
    tensors live wherever the backend wants them (cuda for nccl, cpu for gloo)."""
    import torch
    world = dist.get_world_size()
    gids = to_global_ids(local_ids, base).contiguous()
    B = gids.shape[0]
    all_i = torch.empty((world, B, k), dtype=gids.dtype, device=gids.device)
    all_d = torch.empty((world, B, k), dtype=local_dists.dtype, device=gids.device)
    dist.all_gather_into_tensor(all_i.view(-1), gids.view(-1))
    dist.all_gather_into_tensor(all_d.view(-1), local_dists.contiguous().view(-1))
    return merge_fn(all_i, all_d, k)


class ShardGroup:
    """ctypes mirror of eps_shard_group (include/epsilla_b200.h): the NCCL exchange INSIDE the library.
    `unique_id` (128 bytes from ShardGroup.unique_id() on one rank) must reach every rank by host means."""

    def __init__(self, unique_id, rank, world, device):
        import ctypes as C
        from .lib import check, load_library
        self.L = load_library()
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        check(self.L.eps_shard_group_create(C.byref(h), buf, int(rank), int(world), int(device)))
        self.h, self.rank, self.world = h, rank, world

    @staticmethod
    def unique_id():
        import ctypes as C
        from .lib import check, load_library
        buf = (C.c_char * 128)()
        check(load_library().eps_shard_unique_id(buf))
        return bytes(buf)

    def search(self, index, id_base, d_queries_ptr, nq, k, d_out_ids_ptr, d_out_dists_ptr, sync=True):
        """Every rank: same device-resident query batch, own shard index -> merged GLOBAL top-k on every rank."""
        import ctypes as C
        from .lib import check
        check(self.L.eps_search_batch_sharded(self.h, index.h, int(id_base), C.c_void_p(d_queries_ptr), int(nq), int(k), None, 0,
                                              C.c_void_p(d_out_ids_ptr), C.c_void_p(d_out_dists_ptr), None, int(bool(sync))))

    def close(self):
        if getattr(self, "h", None):
            self.L.eps_shard_group_destroy(self.h)
            self.h = None
