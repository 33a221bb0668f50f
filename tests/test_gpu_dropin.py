"""GPU test of the drop-in boundary: the reference engine's OWN classes (TableSegmentMVP, ANNGraphSegment,
VecSearchExecutor, Expr parser — compiled unmodified) with VecSearchExecutor::{ctor,Search} and
ANNGraphSegment::BuildFromVectorTable replaced at link time by integration/epsilla_b200_dropin.cpp, i.e.
the search and the build run on the B200 through the C ABI while the caller-facing surface is the
reference's.  Replays the reference's golden tests through that surface."""
import os

import numpy as np
import pytest

from helpers import assert_same_results, exact_topk, gen, recall

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "integration", "_build", "libepsilla_ref_b200.so")


@pytest.fixture(scope="module")
def Drop():
    if not os.path.exists(DROPIN):
        pytest.skip("integration/_build/libepsilla_ref_b200.so was not built (needs /root/reference at build time)")
    from oracle.oracle import Ref

    def make(metric, dim, cap, cols=()):
        return Ref(metric, dim, cap, cols, lib_path=DROPIN)
    return make


def test_dense_vector_through_reference_surface(Drop, golden):
    g = golden["dense_vector"]
    for m in ("l2", "ip", "cosine"):
        r = Drop(m, 4, 16, [("ID", "int4")])
        r.set_rows(g["stored_" + m])
        r.set_attr_column("ID", g["ids"])
        r.make_executors(2, 4, 500)  # TableMVP-style: several executors share one device mirror
        ids, ds, _ = r.search(g["query_" + m], 100)
        assert np.array_equal(ids, g["order_" + m])
        assert np.allclose(ds, g["dist_" + m], rtol=1e-4, atol=1e-7)
        ids, ds, _ = r.search(g["query_" + m], 100, "ID <= 2")  # string -> reference parser -> device filter
        assert np.array_equal(ids, g["filter_order_" + m])


def test_halfcircle_through_reference_surface(Drop, golden):
    g = golden["halfcircle"]
    perm = g["perm"]
    r = Drop("cosine", 2, 10000, [("ID", "int4")])
    r.set_rows(g["vectors"])
    r.set_attr_column("ID", perm)
    r.set_row_count(5000)
    r.set_graph(5000, g["offsets"], g["nbrs"].astype(np.int64), int(g["nav"]))
    r.make_executors(1, 4, 500)
    ids, ds, _ = r.search(g["query"], 500)
    assert np.array_equal(perm[ids], np.sort(perm[:5000])[:500])
    r.set_row_count(10000)  # rows appended after the executor was built: tail brute force + merge
    ids, ds, _ = r.search(g["query"], 500)
    assert np.array_equal(perm[ids], np.arange(500))
    r.set_deleted([int(np.nonzero(perm == 3)[0][0])])  # DeleteByPK: deleted rows vanish from results
    ids, ds, _ = r.search(g["query"], 500)
    assert 3 not in perm[ids].tolist() and len(ids) == 500 - 1


def test_gpu_build_through_reference_surface(Drop):
    """ANNGraphSegment::BuildFromVectorTable -> device build -> reference CSR members -> search."""
    n, d, nq = 5000, 32, 32
    X, Q = gen(n, d, 77), gen(nq, d, 78)
    r = Drop("l2", d, n, [("ID", "int4")])
    r.set_rows(X)
    r.set_attr_column("ID", np.arange(n))
    ni, off, nb, nav = r.build(threads=1)
    assert ni == n and off[-1] == len(nb) and nb.min() >= 0 and nb.max() < n
    r.make_executors(4, 1, 500)
    ids, ds, cnt = r.search_batch(Q, 10, "ID >= 100")
    assert np.all(cnt == 10) and np.all(ids >= 100)
    truth = exact_topk(X[100:], Q, 10) + 100
    assert recall(ids, truth, 10) >= 0.97


def test_concurrent_calls_are_coalesced_and_correct(Drop):
    """16 executors hammered by 16 threads (ExecutorPool-style concurrency): single-query Search calls are
    gathered into batched launches by the drop-in's batch former; every caller still gets its own answer."""
    n, d, nq = 20000, 48, 512
    X, Q = gen(n, d, 91), gen(nq, d, 92)
    r = Drop("l2", d, n, [("ID", "int4")])
    r.set_rows(X)
    r.make_executors(16, 4, 500)      # n_indexed = 0 -> exact-scan branch
    ids, ds, cnt = r.search_batch(Q, 10)
    truth = exact_topk(X, Q, 10)
    assert np.all(cnt == 10) and recall(ids, truth, 10) > 0.999
    ids5, _, cnt5 = r.search_batch(Q[:64], 5)   # a different limit goes into its own batch
    assert np.all(cnt5 == 5) and np.array_equal(ids5, ids[:64, :5])


def test_graph_file_round_trip_through_reference_io(Drop, tmp_path):
    """GPU-built graph -> reference SaveANNGraph (ann_graph_<field>.bin) -> reference loader -> GPU search."""
    n, d, nq = 4000, 16, 16
    X, Q = gen(n, d, 61), gen(nq, d, 62)
    r = Drop("l2", d, n, [("ID", "int4")])
    r.set_rows(X)
    ni, off, nb, nav = r.build(threads=1)
    r.make_executors(1, 1, 500)
    before, bd, _ = r.search_batch(Q, 10)
    r.save_graph(str(tmp_path))
    path = tmp_path / "0" / "ann_graph_1.bin"
    raw = np.fromfile(path, dtype=np.int64)       # layout: n, first_id, offsets[n+1], nbrs[E], nav (:171-184)
    assert raw[0] == n and raw[1] == 0 and raw[-1] == nav and len(raw) == 2 + (n + 1) + off[-1] + 1
    assert np.array_equal(raw[2:2 + n + 1], off) and np.array_equal(raw[2 + n + 1:-1], nb)
    r2 = Drop("l2", d, n, [("ID", "int4")])
    r2.set_rows(X)
    r2.load_graph(str(tmp_path))
    r2.make_executors(1, 1, 500)
    after, ad, _ = r2.search_batch(Q, 10)
    assert np.array_equal(before, after) and np.allclose(bd, ad)
