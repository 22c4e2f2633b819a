"""BASELINE.json configs 3 and 5 on one B200 (run under gpurun; writes JSON lines to gpurun_out/).

    python tests/checks/config_runs.py latency   [--rows 10000000 --dim 128 --metric L2 --queries 1000]
        config 3: single-query top-100, one query per call through the HOST entry point (H2D + scan + finalize + D2H +
        synchronise inside every call); p50 / p90 / p99 latency and the algorithmic scan GB/s they imply.
    python tests/checks/config_runs.py streaming [--rows 6250000 --appends 125000 --append-batch 512 --query-batch 256]
        config 5 (one rank's share of 50M rows + 1M appends over 8 ranks): batched appends interleaved with batched
        queries; append rows/s, query QPS and a visibility check (a row appended before a query is found by it).
    python tests/checks/config_runs.py knnjoin   [--rows 2000000 --dim 200 --metric InnerProduct --queries 65536]
        the offline all-pairs job (KnnHelper.findNearestNeighbours): ann_knn_join (one native call, copies overlapped with
        the kernels) against the same job as a loop of blocking ann_query_batch calls; both from host buffers, results equal.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import FuturePool, Metric  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("mode", choices=["latency", "streaming", "config1", "knnjoin"])
ap.add_argument("--rows", type=int, default=None)
ap.add_argument("--dim", type=int, default=None)
ap.add_argument("--metric", default=None)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--appends", type=int, default=125_000)
ap.add_argument("--append-batch", type=int, default=512)
ap.add_argument("--query-batch", type=int, default=256)
a = ap.parse_args()
dev = torch.device("cuda", 0)
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def build(metric, n, d, extra=0):
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0001)
    ix = BruteForceIndex(metric, FuturePool.immediate_pool(), capacity_hint=n + extra)
    for c0 in range(0, n, 1_000_000):
        m = min(1_000_000, n - c0)
        rows = torch.randn((m, d), generator=g, device=dev) / d ** 0.5
        ix.append_batch_device(torch.arange(c0, c0 + m, device=dev, dtype=torch.int64), rows)
    return ix, g


if a.mode == "config1":
    # BASELINE.json configs[0]: Cosine top-100 over 100K x 200, 1K queries -- the reference's own CPU-runnable case.
    # Full parity against the canonical oracle, and the reference-faithful restatement timed on the host cores.
    import oracle

    n, d, nq = a.rows or 100_000, a.dim or 200, a.queries
    metric = Metric.from_string(a.metric or "Cosine")
    rng = np.random.default_rng(0x5EED)
    corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    ids = np.arange(n, dtype=np.int64)
    q = rng.uniform(-1, 1, (nq, d)).astype(np.float32)
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool())
    ix.append_batch(ids, corpus)
    for _ in range(3):
        gi, gd, gc = ix.batch_query_with_distance(q, a.k)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        gi, gd, gc = ix.batch_query_with_distance(q, a.k)
    gpu_s = (time.perf_counter() - t0) / reps
    lat = []
    for i in range(200):
        t0 = time.perf_counter()
        ix.batch_query_with_distance(q[i:i + 1], a.k)
        lat.append((time.perf_counter() - t0) * 1e3)
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, a.k)
    fx = oracle.FaithfulIndex(metric.ordinal, d)
    fx.append(ids, corpus)
    cores = oracle.max_threads()
    t0 = time.perf_counter()
    fi, fd, fc = fx.query(q, a.k, nthreads=1)
    cpu1_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    fx.query(q, a.k, nthreads=cores)
    cpuN_s = time.perf_counter() - t0
    rec = {"config": f"config 1: {metric.name} top-{a.k}, {n}x{d} fp32, {nq} queries",
           "gpu_batch_qps_host_api": nq / gpu_s, "gpu_ms_per_1k_batch": gpu_s * 1e3,
           "gpu_single_query_p50_ms": float(np.median(lat)), "gpu_single_query_p99_ms": float(np.sort(lat)[int(0.99 * len(lat))]),
           "ids_identical_to_oracle": bool((gi == oi).all()), "distance_bits_identical": bool((gd.view(np.uint32) == od.view(np.uint32)).all()),
           "max_rel_dist_err": float(np.max(np.abs(gd - od) / np.maximum(np.abs(od), 1e-30))),
           "faithful_heap_vs_canonical_id_mismatches": int((fi != oi).sum()),
           "cpu_port_qps_1_thread": nq / cpu1_s, f"cpu_port_qps_{cores}_threads": nq / cpuN_s, "cpu_cores": cores,
           "path": {1: "scan", 2: "gemm"}[ix.stat("last_path")]}
    print(json.dumps(rec), flush=True)
    (OUT / "config1.json").write_text(json.dumps(rec, indent=1))
    sys.exit(0 if rec["ids_identical_to_oracle"] else 1)

if a.mode == "latency":
    n, d = a.rows or 10_000_000, a.dim or 128
    metric = Metric.from_string(a.metric or "L2")
    ix, g = build(metric, n, d)
    q = (torch.rand((a.queries + 20, d), generator=g, device=dev) * 2 - 1).cpu().numpy()
    for i in range(20):
        ix.batch_query_with_distance(q[i:i + 1], a.k)
    lat = []
    for i in range(20, 20 + a.queries):
        t0 = time.perf_counter()
        ix.batch_query_with_distance(q[i:i + 1], a.k)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.sort(np.array(lat))
    p = lambda x: float(lat[min(len(lat) - 1, int(x * len(lat)))])
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    bytes_q = n * d * 4
    rec = {"config": f"config 3: single-query {metric.name} top-{a.k}, {n}x{d} fp32, one query per host call", "queries": a.queries,
           "p50_ms": p(0.5), "p90_ms": p(0.9), "p99_ms": p(0.99), "mean_ms": float(lat.mean()), "qps_single_stream": 1e3 / float(lat.mean()),
           "scan_GBps_at_p50": bytes_q / (p(0.5) * 1e-3) / 1e9, "hbm_peak_GBps_measured": hbm,
           "frac_of_measured_hbm_at_p50": bytes_q / (p(0.5) * 1e-3) / 1e9 / hbm, "frac_of_nominal_8TBps_at_p50": bytes_q / (p(0.5) * 1e-3) / 8e12,
           "path": {1: "scan", 2: "gemm"}[ix.stat("last_path")], "includes": "H2D query, scan, finalize, D2H results, stream sync"}
    print(json.dumps(rec), flush=True)
    (OUT / "config3_latency.json").write_text(json.dumps(rec, indent=1))

if a.mode == "streaming":
    n, d = a.rows or 6_250_000, a.dim or 200
    metric = Metric.from_string(a.metric or "InnerProduct")
    ix, g = build(metric, n, d, extra=a.appends + 16)
    nb = a.appends // a.append_batch
    new_rows = (torch.randn((nb * a.append_batch, d), generator=g, device=dev) / d ** 0.5).cpu().numpy()
    queries = (torch.rand((a.query_batch, d), generator=g, device=dev) * 2 - 1).cpu().numpy()
    ix.batch_query_with_distance(queries, a.k)  # warm-up
    t_app = t_q = 0.0
    n_q = 0
    visible_ok = True
    next_id = n
    t_all = time.perf_counter()
    for bi in range(nb):
        rows = new_rows[bi * a.append_batch:(bi + 1) * a.append_batch]
        ids = np.arange(next_id, next_id + a.append_batch, dtype=np.int64)
        t0 = time.perf_counter()
        ix.append_batch(ids, rows)               # host rows -> device matrix + norms/shadow kernel, returns when visible
        t_app += time.perf_counter() - t0
        next_id += a.append_batch
        if bi % 8 == 7:                          # a query batch every 8 append batches
            qb = queries.copy()
            if metric.name != "InnerProduct":
                qb[0] = rows[-1]                 # visibility probe: the row appended just before this query (distance 0)
            else:
                qb[0] = rows[-1] * 64.0          # InnerProduct: make the fresh row the clear winner for its own direction
            t0 = time.perf_counter()
            ids_out, dist_out, _ = ix.batch_query_with_distance(qb, a.k)
            t_q += time.perf_counter() - t0
            n_q += a.query_batch
            visible_ok &= bool(ids_out[0, 0] == ids[-1]) if metric.name != "InnerProduct" else bool(ids[-1] in ids_out[0])
    wall = time.perf_counter() - t_all
    rec = {"config": f"config 5 (one rank's share): {n} preloaded rows x {d}, {nb * a.append_batch} rows appended in batches of "
                     f"{a.append_batch}, {metric.name} top-{a.k} query batches of {a.query_batch} every 8 append batches",
           "append_rows_per_s": nb * a.append_batch / t_app, "append_ms_per_batch": 1e3 * t_app / nb,
           "query_qps": n_q / t_q if t_q else None, "query_ms_per_batch": 1e3 * t_q / max(1, n_q // a.query_batch),
           "wall_s": wall, "final_size": ix.size(), "appended_rows_visible_to_next_query": visible_ok,
           "path": {1: "scan", 2: "gemm"}[ix.stat("last_path")]}
    print(json.dumps(rec), flush=True)
    (OUT / "config5_streaming.json").write_text(json.dumps(rec, indent=1))
    sys.exit(0 if visible_ok else 1)

if a.mode == "knnjoin":
    from the_algorithm_b200.ann.knn import knn_join

    n, d = a.rows or 2_000_000, a.dim or 200
    metric = Metric.from_string(a.metric or "InnerProduct")
    nq = a.queries if a.queries != 1000 else 65536
    rng = np.random.default_rng(1)
    corpus = rng.standard_normal((n, d), dtype=np.float32) / np.float32(d ** 0.5)
    ids = np.arange(n, dtype=np.int64)
    q = rng.random((nq, d), dtype=np.float32) * 2 - 1
    knn_join(q[:64], ids[:4096], corpus[:4096], metric, 10)      # CUDA context, module load, pinned-allocator warm-up
    t_join = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        ji, jd, jc = knn_join(q, ids, corpus, metric, a.k, query_tile=4096)
        t_join = min(t_join, time.perf_counter() - t0)
    t0 = time.perf_counter()
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool(), capacity_hint=n)
    for c0 in range(0, n, 1 << 20):
        ix.append_batch(ids[c0:c0 + (1 << 20)], corpus[c0:c0 + (1 << 20)])
    t_build = time.perf_counter() - t0
    li = np.empty((nq, a.k), np.int64)
    ld = np.empty((nq, a.k), np.float32)
    t_loop = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        for q0 in range(0, nq, 4096):
            i, dd, _ = ix.batch_query_with_distance(q[q0:q0 + 4096], a.k)
            li[q0:q0 + 4096], ld[q0:q0 + 4096] = i, dd
        t_loop = min(t_loop, time.perf_counter() - t0)
    ix.close()
    same = bool((li == ji).all() and (ld.view(np.uint32) == jd.view(np.uint32)).all())
    rec = {"mode": "knnjoin", "rows": n, "dim": d, "metric": metric.name, "queries": nq, "k": a.k,
           "ann_knn_join_s": round(t_join, 3), "blocking_build_s": round(t_build, 3), "blocking_query_loop_s": round(t_loop, 3),
           "join_query_phase_s_estimate": round(t_join - t_build, 3), "identical": same,
           "note": "host numpy buffers in and out (pageable); the join includes building the index from host rows"}
    print(json.dumps(rec), flush=True)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "knnjoin.json").write_text(json.dumps(rec) + "\n")
    sys.exit(0 if same else 1)
