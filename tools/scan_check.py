"""Developer tool: exact-scan step time on the bench table, tensor-core coarse pass with the guard on / off.
  python tools/scan_check.py [rows] [dist ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectordb_b200
from bench import gen_table, gen_queries

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dists = sys.argv[2:] or ["uniform", "cluster"]
dim, nq, k = 768, 1024, 10
dev = torch.device("cuda", 0)
centers = int(os.environ.get("CENTERS", "1024"))
for dist in dists:
    X = gen_table(rows, dim, dist, 42, dev, centers)
    Q = gen_queries(nq, dim, dist, 43, dev, centers)
    ix = vectordb_b200.Index("l2", dim, capacity=rows)
    ix.adopt_device_rows(X.data_ptr(), rows)
    torch.cuda.synchronize()
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int64, device=dev)
    ix.config(512, 512, force_brute=True)
    for mode in os.environ.get("MODES", "bf16,tf32").split(","):
        for guard in [int(x) for x in os.environ.get("GUARDS", "1,0").split(",")]:
            ix.set_coarse(mode); ix.set_coarse_guard(guard)
            for _ in range(int(os.environ.get("WARM", "3"))):
                ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
            ms = 0.0
            reps = int(os.environ.get("REPS", "5"))
            for _ in range(reps):
                st = ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True)
                ms += st["total_ms"]
            print(json.dumps({"dist": dist, "coarse": mode, "guard": guard, "ms_per_step": ms / reps, "kernel_ms": st["kernel_ms"],
                              "launches": st["kernel_launches"], "n_redone": st["n_redone"],
                              "TFLOPs": 2.0 * rows * nq * dim / (ms / reps / 1e3) / 1e12}))
    ix.close(); del X, Q
    torch.cuda.empty_cache()
