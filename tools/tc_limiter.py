"""Developer tool: bracket the limiter of tc_dist_kernel by switching off one engine at a time (EPS_TC_DEBUG bits:
1 = no TMA loads after the first ring fill, 2 = no MMAs, 4 = no epilogue TMEM reads).  Results are garbage in the
debug modes; only the times matter.  One process, the env var is read at every launch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    import torch, vectordb_b200
    rows, dim, nq, k = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (4000000, 768, 1024, 10)))
    coarse = sys.argv[5] if len(sys.argv) > 5 else "bf16"
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    X = torch.rand((rows, dim), generator=g, device=dev)
    Q = torch.rand((nq, dim), generator=g, device=dev)
    ix = vectordb_b200.Index("l2", dim, capacity=rows)
    ix.adopt_device_rows(X.data_ptr(), rows)
    ix.config(512, 512, force_brute=True)
    ix.set_coarse(coarse)
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int64, device=dev)
    out = {"case": [rows, dim, nq, k, coarse], "flops": 2.0 * rows * dim * nq}
    ref = None
    plan = [(0, 0), (1, 0), (2, 0), (4, 0), (5, 0), (6, 0), (0, 0)]
    if os.environ.get("TC_LIMITER_PLAN") == "epi":
        plan = [(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (0, 0), (0, 2)]
    for mode, epi in plan:
        os.environ["EPS_TC_DEBUG"] = str(mode)
        os.environ["EPS_TC_EPI"] = str(epi)
        ms = []
        for _ in range(4):
            st = ix.search_device(Q.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), want_stats=True)
            ms.append(st["kernel_ms"])
        torch.cuda.synchronize()
        if mode == 0:
            ids = oi.cpu()
            if ref is None: ref = ids
            else: out["mode0_ids_equal_first"] = bool((ids == ref).all()) and out.get("mode0_ids_equal_first", True)
        key = {0: "all", 1: "no_tma", 2: "no_mma", 4: "no_epilogue", 5: "mma_only", 6: "tma_only"}[mode] + ("_epi%d" % epi if epi else "")
        while key + "_ms" in out: key += "_again"
        out[key + "_ms"] = round(min(ms[1:]), 3)
        out[key + "_TFs"] = round(out["flops"] / min(ms[1:]) / 1e9, 1)
    os.environ["EPS_TC_DEBUG"] = "0"; os.environ.pop("EPS_TC_EPI", None)
    print(json.dumps(out))

if __name__ == "__main__":
    main()
