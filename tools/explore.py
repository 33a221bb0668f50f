"""Developer tool: build a graph on device and sweep L; prints recall, distance evaluations, kernel time and
the HBM-roofline fraction of graph_search_kernel.  Usage: python tools/explore.py rows dim dist [knn_k] [L,...]"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectordb_b200
from bench import gen_table, gen_queries

rows, dim, dist = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
knn_k = int(sys.argv[4]) if len(sys.argv) > 4 else 64
Ls = [int(x) for x in (sys.argv[5] if len(sys.argv) > 5 else "128,512,2048").split(",")]
exact_below = int(os.environ.get("EXACT_BELOW", "60000"))
nq, k = 1024, 10
dev = torch.device("cuda", 0)
centers = int(os.environ.get("CENTERS", "1024"))
X = gen_table(rows, dim, dist, 42, dev, centers)
Q = gen_queries(nq, dim, dist, 43, dev, centers)
ix = vectordb_b200.Index("l2", dim, capacity=rows)
ix.adopt_device_rows(X.data_ptr(), rows)
torch.cuda.synchronize()  # generators done before the library stream reads
oi = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
oc = torch.empty((nq,), dtype=torch.int64, device=dev)
ix.config(512, 512, force_brute=True)
st = ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True)
truth = oi.cpu().numpy().copy()
print("brute: %.1f ms" % st["kernel_ms"])
t0 = time.perf_counter()
ix.build(rows, knn_k=knn_k, exact_knn_below=exact_below, nnd_iters=int(os.environ.get("NND_ITERS", "12")),
         nnd_sample=int(os.environ.get("NND_S", "32")), out_degree=int(os.environ.get("OUT_DEG", "50")), min_degree=int(os.environ.get("MIN_DEG", "0")))
torch.cuda.synchronize()
print("build: %.2f s" % (time.perf_counter() - t0))
n, off, nb, nav = ix.get_graph()
deg = np.diff(off)
print("edges %d avg deg %.1f max %d nav %d" % (off[-1], deg.mean(), deg.max(), nav))
tunes = [tuple(int(y) for y in x.split("x")) for x in os.environ.get("TUNES", "0x0").split(",")]  # ring x ctas
for L in Ls:
 for W in [int(x) for x in os.environ.get('WIDTHS', '1').split(',')]:
  for (ring, ctas) in tunes:
    ix.config(L, L)
    ix.set_search_width(W)
    ix.set_graph_tuning(ring, ctas)
    for rep in range(2):
        st = ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True)
    g = oi.cpu().numpy()
    rec = np.mean([len(set(g[i]) & set(truth[i])) / k for i in range(nq)])
    nd, ns = st["n_dist"], st["n_seed"]
    byt = (nd - ns) * dim * 4 + st["n_edges"] * 4 + st["n_expand"] * 16 + L * dim * 4 + nq * (dim * 4 + k * 12)
    gbs = byt / (st["kernel_ms"] / 1e3) / 1e9
    print(json.dumps({"L": L, "W": W, "ring": ring, "ctas": ctas, "recall": round(float(rec), 4), "n_dist_per_q": nd / nq, "n_expand_per_q": st["n_expand"] / nq,
                      "kernel_ms": round(st["kernel_ms"], 3), "qps": round(nq / (st["kernel_ms"] / 1e3)), "GBps": round(gbs, 1),
                      "frac_of_6487": round(gbs / 6487.1, 3)}))
