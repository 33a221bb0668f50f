"""Developer tool: build a graph once (saved under /tmp), then search it in a separate process — so that
compute-sanitizer / ncu can wrap the search alone.  Also checks the exact mode (width 1) against the C port.
  python tools/repro.py build rows dim centers
  python tools/repro.py search rows dim centers L width [ring ctas]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectordb_b200
from bench import gen_table, gen_queries

what, rows, dim, centers = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda", 0)
DIST = os.environ.get("DIST", "cluster")
X = gen_table(rows, dim, DIST, 42, dev, centers)
nq, k = int(os.environ.get("NQ", "1024")), 10
Q = gen_queries(nq, dim, DIST, 43, dev, centers)
ix = vectordb_b200.Index("l2", dim, capacity=rows)
ix.adopt_device_rows(X.data_ptr(), rows)
torch.cuda.synchronize()  # generators done before the library stream reads
path = "/tmp/repro_graph_%s_%d_%d_%d.npz" % (DIST, rows, dim, centers)
if what == "build":
    t0 = time.perf_counter()
    ix.build(rows, knn_k=64, nnd_iters=int(os.environ.get("NND_ITERS", "10")))
    n, off, nb, nav = ix.get_graph()
    deg = np.diff(off)
    print("build %.1f s, edges %d, avg deg %.1f, max deg %d" % (time.perf_counter() - t0, off[-1], deg.mean(), deg.max()))
    np.savez(path, off=off, nb=np.asarray(nb, dtype=np.int32), nav=nav)
    sys.exit(0)
g = np.load(path)
ix.set_graph(rows, g["off"], g["nb"].astype(np.int64), int(g["nav"]))
L, width = int(sys.argv[5]), int(sys.argv[6])
ring, ctas = (int(sys.argv[7]), int(sys.argv[8])) if len(sys.argv) > 8 else (0, 0)
oi = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
oc = torch.empty((nq,), dtype=torch.int64, device=dev)
ix.config(512, 512, force_brute=True); ix.set_coarse("fp32")
ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
truth = oi.cpu().numpy().copy()
ix.config(L, L); ix.set_search_width(width); ix.set_graph_tuning(ring, ctas)
st = ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True)
got = oi.cpu().numpy()
rec = np.mean([len(set(got[i]) & set(truth[i])) / k for i in range(nq)])
out = {"L": L, "width": width, "recall": float(rec), "n_dist_per_q": st["n_dist"] / nq, "kernel_ms": st["kernel_ms"]}
if os.environ.get("PORT_CHECK"):
    from oracle.oracle import Port
    Xh = X.cpu().numpy(); Qh = Q[:8].cpu().numpy()
    pids, pds, pcnt, (nd, _) = Port().search_batch(metric="l2", vectors=Xh, queries=Qh, limit=k, n_indexed=rows, offsets=g["off"],
                                                   nbrs=g["nb"].astype(np.int64), nav=int(g["nav"]), L=L)
    out["port_ids_match"] = float((pids == got[:8]).mean())
    out["port_recall"] = float(np.mean([len(set(pids[i]) & set(truth[i])) / k for i in range(8)]))
    out["port_n_dist_per_q"] = nd / 8
print(json.dumps(out))
