"""Runs BASELINE.json configs C3 and C4 once on one B200 and prints one JSON line each (for profiles/).
  C3: 10M x 768 f32 L2, batch = 4096, top-100 + metadata filter (attr = row % 100, "attr < 10")
  C4: 10M x 1536 f32 IP, graph build on device + search (batch 1024, top-10)
Usage: python tools/config_runs.py c3|c4 [rows]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectordb_b200
from bench import gen_table, gen_queries

which = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        st = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, st


if which == "c1":
    # C1: brute-force L2 top-10, 10k x 128 f32, batch = 1, through the host-buffer API, next to the reference CPU path
    from oracle.oracle import Ref
    rows, dim, k = 10_000, 128, 10
    rng = np.random.default_rng(42)
    X = rng.random((rows, dim), dtype=np.float32)
    Q = rng.random((512, dim), dtype=np.float32)
    ix = vectordb_b200.Index("l2", dim, host_vectors=X)
    ix.sync_rows(rows)
    for q in Q[:32]:
        ix.search(q, k, want_stats=False)
    t0 = time.perf_counter()
    got = [ix.search(q, k, want_stats=False)[0][0] for q in Q]
    gpu_qps = len(Q) / (time.perf_counter() - t0)
    out = {"config": "C1 10000x128 L2 brute force, batch=1, top-10 (one eps_search_batch call per query, host buffers)",
           "gpu_qps_sequential_calls": gpu_qps, "gpu_latency_us": 1e6 / gpu_qps}
    try:
        r = Ref("l2", dim, rows, [("ID", "int4")])
        r.set_rows(X)
        for T in (1, 8):
            r.make_executors(1, T, 500)
            for q in Q[:8]:
                r.search(q, k)
            t0 = time.perf_counter()
            ref = [r.search(q, k)[0] for q in Q]
            out["reference_cpu_qps_T%d" % T] = len(Q) / (time.perf_counter() - t0)
        out["ids_identical"] = float(np.mean([np.array_equal(a, b) for a, b in zip(got, ref)]))
    except Exception as e:
        out["reference"] = "unavailable: %r" % (e,)
    print(json.dumps(out))
elif which == "c3":
    dim, nq, k = 768, 4096, 100
    X = gen_table(rows, dim, "uniform", 42, dev)
    Q = gen_queries(nq, dim, "uniform", 43, dev)
    ix = vectordb_b200.Index("l2", dim, capacity=rows)
    ix.adopt_device_rows(X.data_ptr(), rows)
    attr = (np.arange(rows) % 100).astype(np.int32)
    ix.set_attrs(attr.view(np.uint8), 4, rows)
    nodes = np.array([[7, 1, -1, -1, 0, 0, 0, 0], [1, 1, -1, -1, 10, 0, 0, -1], [19, 3, 0, 1, 0, 0, 0, -1]], np.int64)
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int64, device=dev)
    out = {"config": "C3 %dx%d L2 uniform, batch=%d, top-%d, filter attr<10 (10%%)" % (rows, dim, nq, k)}
    ix.config(512, 512, force_brute=True)
    for mode in ("fp32", "bf16"):
        if mode == "fp32" and rows > 2_000_000:
            continue
        ix.set_coarse(mode)
        dt, st = timed(lambda: ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), filter_nodes=nodes, want_stats=True))
        ids = oi.cpu().numpy()
        out["exact_scan_%s" % mode] = {"qps": nq / dt, "ms_per_batch": dt * 1e3, "all_pass_filter": bool(np.all(attr[ids] < 10)),
                                       "full_counts": bool(np.all(oc.cpu().numpy() == k))}
    # cross-check 4 queries against a float64 scan of the passing rows
    sel = torch.arange(0, rows, device=dev)[torch.arange(0, rows, device=dev) % 100 < 10]
    qs = Q[:4].double()
    best = torch.full((4, k), float("inf"), device=dev, dtype=torch.float64); bid = torch.zeros((4, k), dtype=torch.int64, device=dev)
    for r0 in range(0, sel.numel(), 200_000):
        idx = sel[r0:r0 + 200_000]
        xb = X[idx].double()
        dd = (qs * qs).sum(1)[:, None] - 2 * qs @ xb.T + (xb * xb).sum(1)[None, :]
        cd = torch.cat([best, dd], 1); ci = torch.cat([bid, idx[None, :].expand(4, -1)], 1)
        best, s = torch.topk(cd, k, dim=1, largest=False); bid = torch.gather(ci, 1, s)
    out["recall_at_100_vs_fp64_filtered_scan"] = float(np.mean([len(set(bid[i].tolist()) & set(oi[i].tolist())) / k for i in range(4)]))
    print(json.dumps(out))
else:
    dim, nq, k = 1536, 1024, 10
    X = gen_table(rows, dim, "uniform", 42, dev)
    Q = gen_queries(nq, dim, "uniform", 43, dev)
    ix = vectordb_b200.Index("ip", dim, capacity=rows)
    ix.adopt_device_rows(X.data_ptr(), rows)
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int64, device=dev)
    out = {"config": "C4 %dx%d IP uniform, graph build on device + search, batch=%d, top-%d" % (rows, dim, nq, k)}
    ix.config(512, 512, force_brute=True)
    ix.set_coarse("fp32" if rows <= 2_000_000 else "tf32")
    ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
    truth = oi.cpu().numpy().copy()
    ix.set_coarse("bf16")
    dt, st = timed(lambda: ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True))
    g = oi.cpu().numpy()
    out["exact_scan_bf16"] = {"qps": nq / dt, "ms_per_batch": dt * 1e3,
                              "recall_at_10": float(np.mean([len(set(g[i]) & set(truth[i])) / k for i in range(nq)]))}
    t0 = time.perf_counter()
    ix.build(rows, knn_k=64, nnd_iters=10)
    torch.cuda.synchronize()
    out["graph_build_s"] = time.perf_counter() - t0
    n, off, nb, nav = ix.get_graph()
    out["graph"] = {"edges": int(off[-1]), "avg_degree": float(np.diff(off).mean()), "max_degree": int(np.diff(off).max())}
    ix.set_search_width(4)
    for L in (512, 2048):
        ix.config(L, L)
        dt, st = timed(lambda: ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True))
        g = oi.cpu().numpy()
        byt = (st["n_dist"] - st["n_seed"]) * dim * 4 + st["n_edges"] * 4
        out["graph_L%d" % L] = {"qps": nq / dt, "recall_at_10": float(np.mean([len(set(g[i]) & set(truth[i])) / k for i in range(nq)])),
                               "n_dist_per_query": st["n_dist"] / nq, "kernel_ms": st["kernel_ms"],
                               "hbm_GBps": byt / (st["kernel_ms"] / 1e3) / 1e9}
    print(json.dumps(out))
