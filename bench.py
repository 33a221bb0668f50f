#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Epsilla vector-search hot path.

Metric (BASELINE.json): QPS at recall@10 >= 0.99 on 10M x 768 float32, batch = 1024, top-10, synthetic
iid-uniform[0,1) vectors, reported with the HBM roofline fraction of the dominant kernel and
with the reference's own CPU path timed on the same box.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of 1024 queries per GPU.  A step is timed with CUDA
events on the index's launch stream, bracketed by barrier + synchronize, MAX over ranks.  `value` has the
queries already resident in HBM; `e2e` goes through the host-buffer C-ABI call (eps_search_batch: H2D of the
queries and D2H of ids/distances/counts inside the timed region).

Operating point: the engine has the reference's own two search modes — exact scan (BruteForceSearch /
PreFilter path) and graph search at queue length L.  bench sweeps the candidates (graph L in --L-sweep, then
the exact scan), measures recall@10 against exact ground truth, and reports the fastest mode with
recall >= 0.99 as `value`; every candidate is listed under "modes".

Multi-GPU (N > 1): the 10M x 768 table (30.7 GB) fits one GPU, so ranks hold replicas and the query stream is
partitioned (independent units, no data-path collective, scaling "weak": each rank searches its own 1024-query
batch).  --shard-rows instead partitions the ROWS (config C5 shape): every rank searches the same batch over
its shard and the per-shard top-k are exchanged with one NCCL all-gather + k-way merge kernel.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QPS @ recall@10>=0.99, 10Mx768 f32, batch=1024"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=10_000_000)
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--metric", default="l2")
    p.add_argument("--dist", default="uniform", choices=["uniform", "cluster"])
    p.add_argument("--modes", default="graph,brute-bf16,brute-tf32", help="candidate modes: graph, brute-bf16, brute-tf32, brute-fp32")
    p.add_argument("--L-sweep", default="512,2048", help="graph queue lengths to try")
    p.add_argument("--graph-rows-max", type=int, default=int(os.environ.get("EPS_BENCH_GRAPH_ROWS_MAX", "0")),
                   help="build/search the graph only when rows <= this (0 = graph mode off)")
    p.add_argument("--width", type=int, default=4, help="graph expansion width (1 = reference sequential order)")
    p.add_argument("--shard-rows", action="store_true")
    p.add_argument("--recall-target", type=float, default=0.99)
    p.add_argument("--cpu-queries", type=int, default=32)
    p.add_argument("--no-cpu", action="store_true")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------------
# synthetic data (identical bits on every rank / in the reference arm: torch CPU generator is not used for
# the 30 GB table — it is generated on device in 1M-row chunks from a seeded Philox stream)
# ------------------------------------------------------------------------------------------------------
def gen_table(rows, dim, dist, seed, device, centers_n=1024):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.empty((rows, dim), dtype=torch.float32, device=device)
    centers = None
    if dist == "cluster":
        gc = torch.Generator(device=device)
        gc.manual_seed(44)
        centers = torch.rand((centers_n, dim), generator=gc, device=device)
    step = 1_000_000
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        if dist == "uniform":
            X[r0:r1].uniform_(0.0, 1.0, generator=g)
        else:
            lab = torch.randint(0, centers_n, (r1 - r0,), generator=g, device=device)
            X[r0:r1].normal_(0.0, 0.1, generator=g)
            X[r0:r1] += centers[lab]
    return X


def gen_queries(n, dim, dist, seed, device, centers_n=1024):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if dist == "uniform":
        return torch.rand((n, dim), generator=g, device=device)
    gc = torch.Generator(device=device)
    gc.manual_seed(44)
    centers = torch.rand((centers_n, dim), generator=gc, device=device)
    lab = torch.randint(0, centers_n, (n,), generator=g, device=device)
    return centers[lab] + 0.1 * torch.randn((n, dim), generator=g, device=device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = the reference's own sources compiled unmodified)
# ------------------------------------------------------------------------------------------------------
class CpuReference:
    """VecSearchExecutor::Search of the reference on the host cores (oracle/_ref = the reference's own sources
    compiled unmodified; the scalar C port only if that library did not travel).  graph=None -> the brute-force
    branch (no graph: the path the reference takes for an un-indexed table); else (n_indexed, offsets, nbrs, nav)
    -> graph search on the SAME CSR the GPU used.  The table is loaded once; search() is timed per call."""

    def __init__(self, X_host, metric, graph=None, L=500, n_exec=None):
        from oracle import oracle
        self.cores = os.cpu_count() or 1
        self.n, self.d = X_host.shape
        self.kind = "reference" if oracle.have_ref() else "port"
        self.metric, self.graph, self.L, self.X = metric, graph, L, X_host
        if self.kind == "reference":
            r = oracle.Ref(metric, self.d, self.n, [("ID", "int4")])
            r.vectors[:self.n] = X_host
            r.set_row_count(self.n)
            if graph is None:
                self.n_exec = min(n_exec or 16, self.cores)
                self.T = max(1, self.cores // self.n_exec)  # BruteForceSearch parallelises its distance loop with OpenMP
            else:
                r.set_graph(*graph)
                self.n_exec, self.T = self.cores, 1          # throughput-optimal: one executor per core
            r.make_executors(self.n_exec, self.T, L)
            self.r = r
        else:
            self.port = oracle.Port()
            self.n_exec, self.T = 1, 1

    def search(self, Q_host, k):
        Qs = np.ascontiguousarray(Q_host)
        t0 = time.perf_counter()
        if self.kind == "reference":
            ids, ds, cnt = self.r.search_batch(Qs, k)
        else:
            kw = dict(metric=self.metric, vectors=self.X, queries=Qs, limit=k, L=self.L)
            if self.graph is not None:
                kw.update(n_indexed=self.graph[0], offsets=self.graph[1], nbrs=self.graph[2], nav=self.graph[3])
            ids, ds, cnt, _ = self.port.search_batch(**kw)
        dt = time.perf_counter() - t0
        return len(Qs) / dt, ids

    def describe(self, nq):
        return "%d queries of the step's batch over all %d rows, %d executors x %d threads" % (nq, self.n, self.n_exec, self.T)


def run_reference_arm(a):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores on a
    bounded sample of the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    X = gen_table(a.rows, a.dim, a.dist, 42, dev)
    Xh = X.cpu().numpy()
    del X
    nq = max(1, min(a.cpu_queries // 4, a.batch))  # bounded sample per step: the whole run ends within minutes
    ref = CpuReference(Xh, a.metric, None, 500, n_exec=nq)  # one executor per sampled query, all host threads busy
    vals = []
    for s in range(a.warmup + a.steps):
        Q = gen_queries(a.batch, a.dim, a.dist, 43 + s * 64, dev).cpu().numpy()
        qps, _ = ref.search(Q[:nq], a.k)
        if s >= a.warmup:
            vals.append(qps)
    kind, used, sample = ref.kind, ref.n_exec * ref.T, ref.describe(nq)
    v = float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * a.batch / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%dx%d f32 %s %s, batch=%d, top-%d, exact scan (reference BruteForceSearch branch)" % (
            a.rows, a.dim, a.metric, a.dist, a.batch, a.k)},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    import vectordb_b200
    from vectordb_b200.index import merge_shards_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hbm_peak, tf_peak, peak_src = measured_peaks()
    rows = a.rows
    # replicas: same table on every rank; --shard-rows: rank r holds rows [r*rows, (r+1)*rows) of an N*rows table
    X = gen_table(rows, a.dim, a.dist, 42 + (rank if a.shard_rows else 0), dev)
    ix = vectordb_b200.Index(a.metric, a.dim, capacity=rows, device=local)
    ix.adopt_device_rows(X.data_ptr(), rows)
    n_pool = a.warmup + a.steps
    # every rank searches its own batch (replicas) or the same batch (row shards)
    qseed = lambda s: 43 + s * 64 + (0 if a.shard_rows else rank)
    Qpool = [gen_queries(a.batch, a.dim, a.dist, qseed(s), dev) for s in range(n_pool)]
    if a.metric == "cosine":
        X /= X.norm(dim=1, keepdim=True)
        Qpool = [q / q.norm(dim=1, keepdim=True) for q in Qpool]
    out_ids = torch.empty((a.batch, a.k), dtype=torch.int64, device=dev)
    out_d = torch.empty((a.batch, a.k), dtype=torch.float32, device=dev)
    out_c = torch.empty((a.batch,), dtype=torch.int64, device=dev)
    stream = torch.cuda.ExternalStream(ix.stream, device=dev)

    # ---- exact ground truth for recall (the exact-scan mode itself; cross-checked in fp64 on a sample) ----
    ix.config(512, 512, force_brute=True)
    ix.set_coarse("fp32")  # ground truth = the fp32 SIMT exact scan (no tensor-core coarse pass), fp64-checked below
    Qt = Qpool[0]
    ix.search_device(Qt.data_ptr(), a.batch, a.k, out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr())
    truth = out_ids.clone()
    truth_d = out_d.clone()
    chk = 0.0
    with torch.no_grad():
        qs = Qt[:4].double()
        best = torch.full((4, a.k), float("inf"), device=dev, dtype=torch.float64)
        bid = torch.zeros((4, a.k), dtype=torch.int64, device=dev)
        for r0 in range(0, rows, 500_000):
            xb = X[r0:r0 + 500_000].double()
            if a.metric == "l2":
                dd = (qs * qs).sum(1)[:, None] - 2 * qs @ xb.T + (xb * xb).sum(1)[None, :]
            elif a.metric == "ip":
                dd = -(qs @ xb.T)
            else:
                dd = 1 - qs @ xb.T
            cat_d = torch.cat([best, dd], 1)
            cat_i = torch.cat([bid, torch.arange(r0, r0 + xb.shape[0], device=dev)[None, :].expand(4, -1)], 1)
            best, sel = torch.topk(cat_d, a.k, dim=1, largest=False)
            bid = torch.gather(cat_i, 1, sel)
            del xb, dd
        chk = float(np.mean([len(set(bid[i].tolist()) & set(truth[i].tolist())) / a.k for i in range(4)]))
    assert chk >= 0.99, "exact-scan ground truth disagrees with the fp64 check: %.3f" % chk

    def recall_of(ids):
        t, g = truth.cpu().numpy(), ids.cpu().numpy()
        return float(np.mean([len(set(g[i].tolist()) & set(t[i].tolist())) / a.k for i in range(a.batch)]))

    # ---- candidate modes ----
    modes = []
    want = [m.strip() for m in a.modes.split(",") if m.strip()]
    graph_ok = "graph" in want and a.graph_rows_max and rows <= a.graph_rows_max
    build_s = None
    if graph_ok:
        t0 = time.perf_counter()
        ix.build(rows, knn_k=64, nnd_iters=10)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        for L in [int(x) for x in a.L_sweep.split(",")]:
            modes.append(("graph", L, ""))
    for w in want:
        if w.startswith("brute"):
            modes.append(("brute", 0, w.split("-")[1] if "-" in w else "tf32"))
    if not modes:
        modes.append(("brute", 0, "tf32"))

    def set_mode(m):
        if m[0] == "graph":
            ix.config(m[1], m[1], force_brute=False)
            ix.set_search_width(a.width)
        else:
            ix.config(512, 512, force_brute=True)
            ix.set_coarse(m[2])

    def timed_device_steps(m, n_steps, first):
        """Device-resident inputs: CUDA events on the launch stream; max over ranks."""
        set_mode(m)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        stats = []
        barrier()
        for s in range(n_steps):
            q = Qpool[(first + s) % n_pool]
            evs[s][0].record(stream)
            st = ix.search_device(q.data_ptr(), a.batch, a.k, out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr(),
                                  want_stats=True, sync=True)
            evs[s][1].record(stream)
            stats.append(st)
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stats

    report = []
    for m in modes:
        set_mode(m)
        ix.search_device(Qt.data_ptr(), a.batch, a.k, out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr())
        rec = recall_of(out_ids)
        timed_device_steps(m, a.warmup, 0)
        ms, stats = timed_device_steps(m, max(2, min(a.steps, 3)), a.warmup)
        qps = world * a.batch * len(stats) / (ms / 1000.0) if not a.shard_rows else a.batch * len(stats) / (ms / 1000.0)
        report.append({"mode": m[0], "L": m[1], "coarse": m[2], "recall_at_%d" % a.k: rec, "qps_probe": qps,
                       "n_dist_per_query": float(np.mean([s["n_dist"] for s in stats])) / a.batch})
    ok = [r for r in report if r["recall_at_%d" % a.k] >= a.recall_target]
    chosen = max(ok, key=lambda r: r["qps_probe"]) if ok else max(report, key=lambda r: r["recall_at_%d" % a.k])
    mode = (chosen["mode"], chosen["L"], chosen["coarse"])

    # ---- timed region: `value` (inputs resident in HBM) ----
    clocks = ClockSampler(local)
    timed_device_steps(mode, a.warmup, 0)
    clocks.start()
    ms_dev, stats = timed_device_steps(mode, a.steps, a.warmup)
    launches = int(sum(s["kernel_launches"] for s in stats))
    kernel_ms = float(sum(s["kernel_ms"] for s in stats))
    n_dist = float(sum(s["n_dist"] for s in stats))
    n_seed = float(sum(s["n_seed"] for s in stats))
    n_exp = float(sum(s["n_expand"] for s in stats))
    n_edges = float(sum(s["n_edges"] for s in stats))

    # ---- e2e: host buffers through the public C-ABI call, H2D + D2H inside the timed region ----
    set_mode(mode)
    Qhost = [torch.empty((a.batch, a.dim), dtype=torch.float32).pin_memory().copy_(q.cpu()) for q in Qpool]
    e_ids = np.empty((a.batch, a.k), np.int64)
    e_d = np.empty((a.batch, a.k), np.float64)
    e_c = np.empty(a.batch, np.int64)
    import ctypes as C
    from vectordb_b200.lib import check

    def e2e_step(s):
        q = Qhost[s % n_pool]
        check(ix.L.eps_search_batch(ix.h, C.c_void_p(q.data_ptr()), a.batch, a.k, None, 0, e_ids.ctypes.data_as(C.c_void_p),
                                    e_d.ctypes.data_as(C.c_void_p), e_c.ctypes.data_as(C.c_void_p), None))

    for s in range(a.warmup):
        e2e_step(s)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(a.steps):
        e2e_step(a.warmup + s)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    clk = clocks.stop()

    # optional row-shard exchange (all-gather of per-shard top-k + merge kernel), timed separately
    exchange_ms = None
    if a.shard_rows and world > 1:
        from vectordb_b200 import sharded
        mi = torch.empty((a.batch, a.k), dtype=torch.int64, device=dev)
        md = torch.empty((a.batch, a.k), dtype=torch.float32, device=dev)

        def merge_fn(all_i, all_d, kk):
            torch.cuda.synchronize()
            merge_shards_device(local, all_i.data_ptr(), all_d.data_ptr(), world, a.batch, kk, mi.data_ptr(), md.data_ptr())
            return mi, md

        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(a.steps):
            sharded.exchange_and_merge(out_ids, out_d, rank * rows, a.k, dist, merge_fn)
        x1.record()
        barrier()
        exchange_ms = x0.elapsed_time(x1) / a.steps
        # correctness of the exchange: merged results of the last timed step's batch vs merged exact ground truth
        set_mode(mode)
        ix.search_device(Qt.data_ptr(), a.batch, a.k, out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr())
        gi, gd = sharded.exchange_and_merge(out_ids, out_d, rank * rows, a.k, dist, merge_fn)
        got = gi.cpu().numpy().copy()
        ti, td = sharded.exchange_and_merge(truth, truth_d, rank * rows, a.k, dist, merge_fn)
        want = ti.cpu().numpy()
        merged_recall = float(np.mean([len(set(got[i].tolist()) & set(want[i].tolist())) / a.k for i in range(a.batch)]))

    units = a.batch * a.steps * (1 if a.shard_rows else world)
    value = units / ((ms_dev + (exchange_ms or 0.0) * a.steps) / 1000.0)
    e2e_value = units / ((ms_e2e + (exchange_ms or 0.0) * a.steps) / 1000.0)

    # ---- roofline of the dominant kernel (algorithmic bytes per SURVEY.md §8d / DESIGN.md) ----
    if mode[0] == "brute":
        bytes_alg = a.steps * (rows * a.dim * 4.0 + a.batch * a.dim * 4.0 + a.batch * a.k * 12.0)
        kernel_name = "bf_dist_tile_kernel + bf_select_kernel"
    else:
        deg_bytes = 4.0  # int32 neighbour ids on device
        bytes_alg = (n_dist - n_seed) * a.dim * 4.0 + n_edges * deg_bytes + n_exp * 16.0 + \
            a.steps * (mode[1] * a.dim * 4.0) + a.steps * a.batch * (a.dim * 4.0 + a.k * 12.0)
        kernel_name = "graph_search_kernel"
    if mode[0] == "brute" and mode[2] != "fp32" and not os.environ.get("EPS_NO_TC"):
        # exact scan at B=1024 is a dense contraction (512 flop/B): tcgen05 kind::tf32 coarse pass + fp32 re-score.
        # Roofline = tensor pipe.  TF32 runs at half the bf16 rate on tcgen05, so peak = measured bf16 / 2.
        flop = a.steps * rows * float(a.batch) * a.dim * 2.0
        ach = flop / (kernel_ms / 1000.0) / 1e12 if kernel_ms > 0 else 0.0
        peak = tf_peak / 2.0 if mode[2] == "tf32" else tf_peak
        # DRAM bytes of ONE launch of the dominant kernel from the committed ncu capture of this workload
        # (profiles/r01_ncu_tc_dist_fused_bf16_10Mx768.md: 4.04 GB read + 6.5 MB written for a 2.63 M-row launch,
        # i.e. exactly the bf16 rows once); only quoted for the configuration it was captured on.
        traffic = 4.039187e9 + 6.485504e6 if (mode[2] == "bf16" and rows == 10_000_000 and a.dim == 768 and a.batch == 1024) else None
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "tc_dist_kernel (tcgen05 kind::%s, fused threshold select) + bf_select_kernel + rescore_kernel" % (
                    "tf32" if mode[2] == "tf32" else "f16/bf16"),
                "peak_source": "%s bf16 sustained (%.0f TF/s)%s" % (peak_src, tf_peak, " / 2 for TF32" if mode[2] == "tf32" else ""),
                "kernel_ms_per_step": kernel_ms / a.steps,
                "hbm_algorithmic_GBps": bytes_alg / (kernel_ms / 1000.0) / 1e9 if kernel_ms > 0 else 0.0}
    else:
        achieved = bytes_alg / (kernel_ms / 1000.0) / 1e9 if kernel_ms > 0 else 0.0
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": kernel_name, "peak_source": peak_src, "kernel_ms_per_step": kernel_ms / a.steps}
        if mode[0] == "brute":
            flop = a.steps * rows * float(a.batch) * a.dim * (3.0 if a.metric == "l2" else 2.0)
            roof["fp32_simt_tflops"] = flop / (kernel_ms / 1000.0) / 1e12 if kernel_ms > 0 else 0.0

    out = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%dx%d f32 %s iid-%s (seed 42), batch=%d, top-%d, mode=%s%s; %s" % (
            rows, a.dim, a.metric, a.dist, a.batch, a.k, mode[0], (" L=%d" % mode[1]) if mode[0] == "graph" else " (exact scan: tcgen05 %s coarse pass + fp32 re-score)" % mode[2],
            "row shards + NCCL all-gather" if a.shard_rows else "replicated table, query stream partitioned over ranks"),
            "recall_at_%d" % a.k: chosen["recall_at_%d" % a.k], "l2_flush": "inputs (%.1f GB table) larger than L2" % (
                rows * a.dim * 4 / 1e9), "fp64_groundtruth_check": chk},
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": a.batch * a.dim * 4,
                "d2h_bytes_per_step": a.batch * a.k * 12 + a.batch * 8},
        "gpu_launches": launches, "clocks": clk, "roofline": roof, "modes": report,
    }
    if build_s is not None:
        out["graph_build_s"] = build_s
    if exchange_ms is not None:
        out["exchange_ms_per_step"] = exchange_ms
        out["merged_recall_at_%d" % a.k] = merged_recall

    # ---- the reference's CPU path on this box's host cores (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not a.no_cpu:
        try:
            Xh = X.cpu().numpy()
            Qh = Qpool[0].cpu().numpy()
            graph = None
            L = 500
            if mode[0] == "graph":
                graph = ix.get_graph()
                L = mode[1]
            nqc = min(a.cpu_queries, a.batch)
            ref = CpuReference(Xh, a.metric, graph, L)
            qps, cids = ref.search(Qh[:nqc], a.k)
            cores, kind, sample = ref.n_exec * ref.T, ref.kind, ref.describe(nqc)
            agree = float(np.mean([len(set(cids[i].tolist()) & set(truth[i].tolist())) / a.k for i in range(len(cids))]))
            out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample,
                                   "ids_agree_with_gpu": agree}
        except Exception as e:  # the bench line must still print
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": "failed: %r" % (e,)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
