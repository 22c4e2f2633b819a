#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                      (default arm: the CUDA engine)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     (reference arm: CPU restatement, rank 0 only)

metric  : exact top-100 queries/sec over a 10M x 200 fp32 corpus (BASELINE.json `metric`, configs[1]:
          InnerProduct, query batch 4096, tcgen05 GEMM path)
step    : one batch of 4096 synthetic queries through the whole query path (prep -> GEMM filter chunks ->
          compaction -> exact fp64 finalize; with N > 1 also the exchange + merge of the per-rank top-k lists)
value   : whole-job queries/s with queries and outputs resident in HBM (CUDA events on the launching stream,
          barrier + synchronize on both sides, max over ranks)
e2e     : the same through the host-buffer C-ABI call ann_query_batch (pinned host queries in, host results out,
          H2D / D2H inside the timed region)
N > 1   : STRONG scaling on the same 10M-row corpus -- rows sharded contiguously across ranks, queries replicated.
          Every rank first publishes k bounds per query from a short seed launch over its shard; after one barrier each
          rank filters its rows against the k-th best bound of ALL ranks (ann_query_seed_device / _finish_device), then
          the exchange step: the fused exchange+merge kernel over NVLink peer memory (each rank pulls, merges and
          pushes 1/N of the batch; ann/exchange.py), or NCCL all-gather + merge kernel if peers cannot be mapped.
roofline: dominant kernel = gemm_filter (tensor bound); achieved = 2*N_local*d*B flop / its CUDA-event time.
cpu_baseline / --impl reference: the reference-faithful C restatement (oracle/oracle.c: linked list of heap rows,
          Scala PriorityQueue mechanics) on the host cores, one query per thread, bounded sample of the same workload.
          The Scala original cannot run here (no JVM, unshipped deps) -- kind = "port".
Synthetic data: rows ~ N(0, 1/d) (TwHIN-like dense, unnormalised), queries ~ U[-1, 1) (Warmup.scala:46-47).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

UNIT = "queries/s"


def metric_name(a) -> str:
    """BASELINE.json's metric for the default arguments; the same sentence with the actual shape otherwise."""
    rows = f"{a.rows // 1_000_000}M" if a.rows % 1_000_000 == 0 else str(a.rows)
    return f"exact top-{a.k} queries/sec, {rows}x{a.dim} fp32 corpus, {a.metric}, query batch {a.batch}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="total corpus rows (sharded across ranks)")
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--metric", default="InnerProduct")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 streaming scan, 2 tensor-core GEMM filter")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=0, help="queries in the CPU sample (0 = one per host thread)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra block (single-query scan, configs[2], configs[3])")
    ap.add_argument("--lat-queries", type=int, default=400, help="sequential single queries in each latency run of the extra block")
    ap.add_argument("--config4-rows", type=int, default=100_000_000,
                    help="total rows of the configs[3] sub-run (Cosine, batch 1024), sharded over the ranks; 0 = skip")
    ap.add_argument("--config5-rows", type=int, default=50_000_000,
                    help="rows pre-loaded in the configs[4] streaming sub-run (sharded over the ranks); 0 = skip")
    ap.add_argument("--config5-appends", type=int, default=1_000_000)
    ap.add_argument("--seed-rows", type=int, default=0, help="rows of the threshold-seeding launch per shard (0 = library default 65536)")
    ap.add_argument("--full-h2d", action="store_true", help="N > 1, e2e leg: every rank copies the WHOLE query batch from the host (A/B)")
    ap.add_argument("--pull-bounds", action="store_true", help="N > 1: consumers pull the peers' bound arrays (A/B of push delivery)")
    ap.add_argument("--full-seeds", action="store_true",
                    help="N > 1: every rank seeds every query and delivers k bounds per query (A/B of sliced seeding)")
    ap.add_argument("--one-round", action="store_true",
                    help="N > 1: seed round only (A/B of the second cross-shard round that shares the k best bounds)")
    ap.add_argument("--no-share-seeds", action="store_true",
                    help="N > 1: every rank filters against its own thresholds only (A/B of the shared seed bounds)")
    return ap.parse_args()


def workload_name(a) -> str:
    """Which BASELINE.json config the arguments are (by shape), or "custom"."""
    shape = (a.rows, a.dim, a.batch, a.k, a.metric)
    if shape == (10_000_000, 200, 4096, 100, "InnerProduct"):
        tag = "configs[1]"
    elif shape == (10_000_000, 128, 1, 100, "L2"):
        tag = "configs[2]"
    elif shape == (100_000_000, 200, 1024, 100, "Cosine"):
        tag = "configs[3]"
    elif shape == (100_000, 200, 1000, 100, "Cosine"):
        tag = "configs[0]"
    else:
        tag = "custom (not a BASELINE.json config)"
    return f"{tag}: exact {a.metric} top-{a.k} over {a.rows}x{a.dim} fp32 corpus, query batch {a.batch}"


def config_dict(a, n_gpus, route=""):
    how = {"fused": "three-phase sharded query: shared seed thresholds, tensor-core filter, the k best bounds of every rank shared "
                    "after the last chunk so that each rank rescores only its share of the global survivors; then each rank "
                    "pulls and merges ITS 1/N slice of the batch over NVLink peer memory (answer stays partitioned, no push)",
           "allgather": "NCCL all-gather of local top-k + merge kernel"}.get(route or "fused", route)
    return {
        "workload": workload_name(a),
        "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k, "distance": a.metric,
        "sharding": "single GPU" if n_gpus == 1 else f"rows sharded contiguously over {n_gpus} ranks, queries replicated, " + how,
        "l2_flush": "none needed: every step streams the corpus shadow (>= 4 GB per 10M rows) through a 126 MB L2",
    }


# ------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline_run(corpus_np, ids_np, queries_np, metric_ord: int, k: int, steps: int, warmup: int, keep: bool = False):
    """Times the reference-faithful restatement (oracle.FaithfulIndex) with one whole query per host thread."""
    import oracle

    cores = oracle.max_threads()
    ix = oracle.FaithfulIndex(metric_ord, corpus_np.shape[1])
    t0 = time.time()
    chunk = 1_000_000
    for c0 in range(0, corpus_np.shape[0], chunk):
        ix.append(ids_np[c0:c0 + chunk], corpus_np[c0:c0 + chunk])
    build_s = time.time() - t0
    nq = queries_np.shape[0]
    per_step = []
    for s in range(warmup + steps):
        t0 = time.time()
        answers = ix.query(queries_np, k, nthreads=cores)
        dt = time.time() - t0
        if s >= warmup:
            per_step.append(dt)
    ix.close()
    tot = sum(per_step)
    return {"value": nq * len(per_step) / tot, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nq} queries per step x {len(per_step)} step(s) against the full {corpus_np.shape[0]}x{corpus_np.shape[1]} "
                      f"corpus, one query per thread; linked-list build {build_s:.1f}s not timed",
            "ms_per_step": 1000.0 * tot / len(per_step), "answers": answers if keep else None}


def gen_host_data(a, device_ok: bool):
    """Synthetic corpus/queries on the host for the CPU arm (generated on the GPU when present, for speed)."""
    import numpy as np
    import torch

    if device_ok:
        dev = torch.device("cuda", 0)
        g = torch.Generator(device=dev)
        g.manual_seed(0x5EED0001)
        parts = []
        for c0 in range(0, a.rows, 1_000_000):
            m = min(1_000_000, a.rows - c0)
            parts.append((torch.randn((m, a.dim), generator=g, device=dev) / a.dim ** 0.5).cpu())
        corpus = torch.cat(parts).numpy()
        g.manual_seed(0x5EED0002)
        q = (torch.rand((a.batch, a.dim), generator=g, device=dev) * 2 - 1).cpu().numpy()
    else:
        g = torch.Generator()
        g.manual_seed(0x5EED0001)
        corpus = (torch.randn((a.rows, a.dim), generator=g) / a.dim ** 0.5).numpy()
        g.manual_seed(0x5EED0002)
        q = (torch.rand((a.batch, a.dim), generator=g) * 2 - 1).numpy()
    return corpus, np.arange(a.rows, dtype=np.int64), q


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path (restated, see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    import torch

    from oracle import METRIC_BY_NAME
    cores = oracle.max_threads()
    corpus, ids, q = gen_host_data(a, torch.cuda.is_available())
    nq = a.cpu_queries or cores
    res = cpu_baseline_run(corpus, ids, q[:nq], METRIC_BY_NAME[a.metric], a.k, a.steps, a.warmup)
    line = {
        "impl": "reference", "metric": metric_name(a), "value": res["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "dtype_detail": "fp32 embeddings, fp64 accumulation, one rounding to fp32",
        "data": "synthetic", "config": config_dict(a, a.gpus),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference arm = C restatement of BruteForceIndex.scala:66-91 (JVM + unshipped deps cannot run here)",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
def result_digest(ids_np, dist_np, cnt_np) -> str:
    """sha256 over the whole batch's neighbour ids, distance bits and counts: equal digests = bit-identical answers."""
    import hashlib

    import numpy as np

    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ids_np, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(dist_np, dtype=np.float32).view(np.uint32).tobytes())
    h.update(np.ascontiguousarray(cnt_np, dtype=np.int32).tobytes())
    return h.hexdigest()[:32]


def parity_against(got, want, tie_tolerant=False):
    """Compare GPU rows with oracle rows for the same queries.  `tie_tolerant`: ids may be permuted inside a run of exactly
    equal distances (the reference's heap order is history-dependent there), and the last run may differ in
    membership when the tie straddles rank k."""
    import numpy as np

    gi, gd, gc = got
    wi, wd, wc = want
    dist_ok = bool((gd.view(np.uint32) == wd.view(np.uint32)).all()) and bool((gc == wc).all())
    ids_ok = bool((gi == wi).all())
    tie_positions = 0
    if tie_tolerant and not ids_ok and dist_ok:
        ok = True
        for r in range(gi.shape[0]):
            n = int(gc[r])
            j = 0
            while j < n:
                e = j
                while e + 1 < n and gd[r, e + 1].view(np.uint32) == gd[r, j].view(np.uint32):
                    e += 1
                if not (gi[r, j:e + 1] == wi[r, j:e + 1]).all():
                    tie_positions += e + 1 - j
                    if e - j == 0 and e != n - 1:       # a difference outside any tie run
                        ok = False
                    elif e != n - 1 and sorted(gi[r, j:e + 1].tolist()) != sorted(wi[r, j:e + 1].tolist()):
                        ok = False
                j = e + 1
        return dist_ok, ok, tie_positions
    return dist_ok, ids_ok, tie_positions


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import _pkg
    _pkg.load()
    from the_algorithm_b200.ann.brute_force import BruteForceIndex
    from the_algorithm_b200.ann.common import FuturePool, Metric
    from the_algorithm_b200.ann.distributed import ShardedBruteForceIndex, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def build_shard(metric, n, d, seed):
        """This rank's contiguous row range of the n x d corpus; every rank draws the same generator stream."""
        lo, hi = shard_range(n, world, rank)
        ix_ = BruteForceIndex(metric, FuturePool.immediate_pool(), device=local_rank, capacity_hint=hi - lo)
        g_ = torch.Generator(device=dev)
        g_.manual_seed(seed)
        for c0 in range(0, n, 1_000_000):
            m = min(1_000_000, n - c0)
            rows = torch.randn((m, d), generator=g_, device=dev) / d ** 0.5
            s_, e_ = max(c0, lo), min(c0 + m, hi)
            if s_ < e_:
                ix_.append_batch_device(torch.arange(s_, e_, device=dev, dtype=torch.int64), rows[s_ - c0:e_ - c0].contiguous())
            del rows
        return ix_, hi - lo

    def gather_rows(t_ids, t_dist, t_cnt):
        """Whole-batch answer on rank 0 (numpy) from every rank's slice -- outside every timed region."""
        part = (t_ids.cpu().numpy(), t_dist.cpu().numpy(), t_cnt.cpu().numpy())
        if world == 1:
            return part
        parts = [None] * world
        dist.all_gather_object(parts, part)
        return tuple(np.concatenate([p_[j] for p_ in parts]) for j in range(3))

    metric = Metric.from_string(a.metric)
    n, d, b, k = a.rows, a.dim, a.batch, a.k

    # ---- build the shard (not timed) ----
    ix, n_local = build_shard(metric, n, d, 0x5EED0001)
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0002)
    q_dev = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
    if a.path:
        ix.set_option("path", a.path)
    if a.seed_rows:
        ix.set_option("gemm_seed_rows", a.seed_rows)

    out_ids = torch.empty((b, k), dtype=torch.int64, device=dev)
    out_dist = torch.empty((b, k), dtype=torch.float32, device=dev)
    out_cnt = torch.empty((b,), dtype=torch.int32, device=dev)
    sx = ShardedBruteForceIndex(ix, device=dev, share_seeds=not a.no_share_seeds, two_round=not a.one_round, push=not a.pull_bounds,
                                sliced_seeds=not a.full_seeds) if world > 1 else None

    def step_device(queries):
        if world == 1:
            ix.query_batch_device(queries, k, out_ids, out_dist, out_cnt, stream.cuda_stream)
            return out_ids, out_dist, out_cnt
        # local three-phase query + pull/merge of THIS rank's 1/N slice of the batch (the answer stays partitioned)
        return sx.batch_query_device(queries, k, stream.cuda_stream, deliver="slice")

    # ---- warm-up, then the device-resident measurement ----
    for _ in range(max(a.warmup, 3)):
        res = step_device(q_dev)
    barrier()
    ix.raise_pending_error()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ix.stat("launches")
    ix.set_option("timing", 1)
    ms_total = timed(lambda: step_device(q_dev), a.steps)
    kernel_us = ix.stat("kernel_us")
    kernel_n = ix.stat("kernel_launches_timed")
    ix.set_option("timing", 0)
    launches = ix.stat("launches") - launches0 + (a.steps if world > 1 else 0)
    ix.raise_pending_error()
    last_path = ix.stat("last_path")
    ms_step = ms_total / a.steps
    value = b / (ms_step * 1e-3)

    # ---- where a step goes: CUDA events around EVERY kernel of the query path (a separate short loop: the extra event
    # records cost a little, so this is not the timed region above); rank 0's view
    bsteps = 5
    ix.set_option("timing", 2)
    ms_b = timed(lambda: step_device(q_dev), bsteps) / bsteps
    breakdown = {"ms_per_step_with_event_records": ms_b}
    acc_ms = 0.0
    for name in ("main", "prep", "compact", "seed_merge", "finalize"):
        v = ix.stat("us_" + name) / 1e3 / bsteps
        breakdown[{"main": "scan_or_gemm_filter"}.get(name, name) + "_ms"] = v
        acc_ms += v
    breakdown["exchange_barriers_launch_gaps_ms"] = ms_b - acc_ms
    ix.set_option("timing", 0)

    # ---- end to end through the host-buffer C-ABI call: pinned queries in, host results out ----
    q_pin = q_dev.cpu().pin_memory()
    q0, q1 = (0, b) if world == 1 else sx.slice_range(b)
    h_ids = torch.empty((q1 - q0, k), dtype=torch.int64).pin_memory()
    h_dist = torch.empty((q1 - q0, k), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((q1 - q0,), dtype=torch.int32).pin_memory()
    q_np = q_pin.numpy()
    out_np = (h_ids.numpy(), h_dist.numpy(), h_cnt.numpy())      # pinned result buffers handed to the C ABI

    def step_e2e():
        if world == 1:
            return ix.batch_query_with_distance(q_np, k, out=out_np)   # ann_query_batch: H2D, query path, D2H inside the call
        # N ranks: rows are sharded, queries replicated on the devices -- but the batch crosses PCIe only once: every rank copies
        # ITS 1/N slice from pinned host memory, pushes it to the peers over NVLink (one barrier), and keeps its slice of the answer
        if sx.route == "fused" and not a.full_h2d:
            qd = sx.gather_queries(q_pin[q0:q1], b, k, stream.cuda_stream)
        else:
            qd = q_pin.to(dev, non_blocking=True)
        oi, od, oc = step_device(qd)
        h_ids.copy_(oi, non_blocking=True)
        h_dist.copy_(od, non_blocking=True)
        h_cnt.copy_(oc, non_blocking=True)
        stream.synchronize()
        return h_ids, h_dist, h_cnt

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / a.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the answer itself: digest of the whole batch (must be the same at every N), no flagged rows ----
    full = gather_rows(torch.from_numpy(out_np[0]), torch.from_numpy(out_np[1]), torch.from_numpy(out_np[2]))
    digest = result_digest(*full)
    flagged_rows = int((full[2] < 0).sum())

    # ---- roofline of the dominant kernel (CUDA events around each launch, on the launching stream) ----
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    traffic, traffic_detail = None, None
    try:   # DRAM bytes of the same kernel from the committed ncu --set full capture of this command (profiles/)
        tj = json.loads(next(p_ for p_ in (ROOT / "profiles" / "traffic.json", ROOT / "profiles" / "r01_traffic.json") if p_.exists()).read_text())
        if (n, d, b, k, world, a.metric) == (10_000_000, 200, 4096, 100, 1, "InnerProduct"):
            traffic = float(tj["gemm_filter_dram_bytes_per_step"])     # dram read + write, summed over the step's launches
            traffic_detail = {"unit": "bytes per step (the kernel's launches of one batch), like `achieved`",
                              "operand_bytes_per_step": tj["algorithmic_shadow_bytes_per_step"], "source": tj["source"]}
    except Exception:
        pass
    if last_path == 2:
        flops_step = 2.0 * n_local * d * b
        achieved = flops_step * a.steps / (kernel_us * 1e-6) / 1e12 if kernel_us else None
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        roofline = {"bound": "tensor", "kernel": "gemm_filter_kernel<2> (tcgen05.mma kind::f16, bf16 in / fp32 accumulate)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks else "fallback 1.4 PFLOP/s sustained",
                    "frac_of_burst": (achieved / float(peaks["bf16_tflops"])) if achieved and peaks.get("bf16_tflops") else None,
                    "traffic": traffic, "traffic_detail": traffic_detail, "kernel_ms_per_step": kernel_us / 1e3 / a.steps,
                    "launches_per_step": kernel_n / a.steps,
                    "algorithmic_flops_per_step": flops_step}
    else:
        bytes_step = float(n_local) * d * 4 * ((b + 7) // 8)
        achieved = bytes_step * a.steps / (kernel_us * 1e-6) / 1e9 if kernel_us else None
        peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                    "kernel_ms_per_step": kernel_us / 1e3 / a.steps, "launches_per_step": kernel_n / a.steps}

    # ---- extra, same process: the HBM-bound single-query scan on this corpus, config 3 and config 4 ----
    extra = {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def latency_run(index, queries_np, nq_lat, scan_bytes_per_query):
        """Sequential single queries through the host call (H2D + kernels + D2H per call): p50 / p99, and the kernel-only
        scan bandwidth from the CUDA events around the scan launches."""
        index.set_option("path", 1)
        for i in range(5):
            index.batch_query_with_distance(queries_np[i:i + 1], k)
        index.set_option("timing", 1)
        lat = []
        for i in range(nq_lat):
            t_ = time.perf_counter()
            index.batch_query_with_distance(queries_np[i % len(queries_np)][None, :], k)
            lat.append((time.perf_counter() - t_) * 1e3)
        kus = index.stat("kernel_us")
        index.set_option("timing", 0)
        index.set_option("path", 0)
        lat.sort()
        p50, p99 = lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))]
        gbs_kernel = scan_bytes_per_query * nq_lat / (kus * 1e-6) / 1e9 if kus else None
        return {"queries": nq_lat, "p50_ms": p50, "p99_ms": p99, "qps_single_stream": 1e3 / (sum(lat) / len(lat)),
                "scan_GBps_kernel": gbs_kernel, "scan_GBps_per_host_call_p50": scan_bytes_per_query / (p50 * 1e-3) / 1e9,
                "hbm_peak_GBps": hbm_peak, "frac_of_hbm_peak_kernel": gbs_kernel / hbm_peak if gbs_kernel else None,
                "frac_of_hbm_peak_per_host_call_p50": scan_bytes_per_query / (p50 * 1e-3) / 1e9 / hbm_peak,
                "algorithmic_bytes_per_query": scan_bytes_per_query}

    if not a.no_extra and world == 1 and last_path == 2:
        try:   # BASELINE.json's metric: "scan GB/s vs HBM peak" -- B = 1 on the headline corpus, fp32 rows streamed from HBM
            r_ = latency_run(ix, q_np, a.lat_queries, float(n) * d * 4 + (float(n) * 4 if a.metric == "Cosine" else 0.0))
            r_["workload"] = f"single-query {a.metric} top-{k} over {n}x{d} fp32 (streaming scan), sequential host calls"
            extra["scan"] = r_
        except Exception as e:
            extra["scan"] = {"error": repr(e)}
    if not a.no_extra and world == 1 and last_path == 2:
        try:   # online shape: 64 host threads, one vector per call, against the same handle (library-side load generator)
            ix.loadtest(q_np[:1024], k, threads=64, calls_per_thread=20)      # untimed: thread start-up, first lone-caller batches
            lt = ix.loadtest(q_np[:1024], k, threads=64, calls_per_thread=400, expect_ids=full[0][:1024])
            lt["workload"] = (f"64 host threads x 400 one-vector ann_query_batch calls (after 64 x 20 untimed), {a.metric} top-{k} over {n}x{d}; concurrent calls "
                              "are combined into device batches by the library's micro-batcher (QueryIndexThriftController.scala:39-90 shape)")
            extra["online_single_vector"] = lt
        except Exception as e:
            extra["online_single_vector"] = {"error": repr(e)}
    ix_closed = False
    if not a.no_extra and world == 1:
        try:   # configs[2]: single-query L2 top-100 over 10M x 128, batch 1, p50/p99 latency
            ix3, _ = build_shard(Metric.from_string("L2"), 10_000_000, 128, 0x5EED0003)
            g.manual_seed(0x5EED0004)
            q3 = (torch.rand((256, 128), generator=g, device=dev) * 2 - 1).cpu().numpy()
            r_ = latency_run(ix3, q3, a.lat_queries, 10_000_000.0 * 128 * 4)
            r_["workload"] = "configs[2]: single-query L2 top-100 over 10000000x128 fp32, batch 1, sequential host calls"
            i3, d3, c3 = ix3.batch_query_with_distance(q3[:64], k)
            r_["result_digest"] = result_digest(i3, d3, c3)
            extra["config3"] = r_
            ix3.close()
        except Exception as e:
            extra["config3"] = {"error": repr(e)}
    if not a.no_extra and a.config4_rows > 0:
        try:   # configs[3]: Cosine top-100 over 100M x 200 row-sharded over the N ranks, batch 1024
            ix.close()
            ix_closed = True
            torch.cuda.empty_cache()
            n4, b4 = a.config4_rows, 1024
            ix4, n4_local = build_shard(Metric.from_string("Cosine"), n4, d, 0x5EED0005)
            g.manual_seed(0x5EED0006)
            q4 = (torch.rand((b4, d), generator=g, device=dev) * 2 - 1).contiguous()
            o4 = (torch.empty((b4, k), dtype=torch.int64, device=dev), torch.empty((b4, k), dtype=torch.float32, device=dev),
                  torch.empty((b4,), dtype=torch.int32, device=dev))
            sx4 = ShardedBruteForceIndex(ix4, device=dev, share_seeds=not a.no_share_seeds, two_round=not a.one_round, push=not a.pull_bounds,
                                sliced_seeds=not a.full_seeds) if world > 1 else None

            def step4():
                if world == 1:
                    ix4.query_batch_device(q4, k, o4[0], o4[1], o4[2], stream.cuda_stream)
                    return o4
                return sx4.batch_query_device(q4, k, stream.cuda_stream, deliver="slice")

            for _ in range(3):
                r4 = step4()
            barrier()
            ix4.raise_pending_error()
            ix4.set_option("timing", 1)
            steps4 = max(3, min(a.steps, 10))
            ms4 = timed(step4, steps4) / steps4
            kus4 = ix4.stat("kernel_us")
            ix4.set_option("timing", 0)
            full4 = gather_rows(*r4)
            tf4 = 2.0 * n4_local * d * b4 * steps4 / (kus4 * 1e-6) / 1e12 if kus4 else None
            extra["config4"] = {"workload": f"configs[3]: Cosine top-{k} over {n4}x{d} fp32 row-sharded over {world} rank(s), batch {b4}",
                                "value": b4 / (ms4 * 1e-3), "unit": UNIT, "ms_per_step": ms4, "steps": steps4, "n_gpus": world,
                                "rows_per_rank": n4_local, "kernel_TFLOPs_per_rank": tf4,
                                "result_digest": result_digest(*full4), "flagged_rows": int((full4[2] < 0).sum())}
            ix4.close()
        except Exception as e:
            extra["config4"] = {"error": repr(e)}

    if not a.no_extra and a.config5_rows > 0:
        try:   # configs[4]: streaming Appendable -- 1M rows appended in host batches, a 256-query batch after every append batch
            if not ix_closed:
                ix.close()
                ix_closed = True
            torch.cuda.empty_cache()
            n5, ab5, qb5, total5 = a.config5_rows, 4096, 256, a.config5_appends
            lo5, hi5 = shard_range(n5, world, rank)
            # capacity_hint = 0: the storage grows in place while the rows arrive (virtual ranges + mapped chunks, no copies)
            ix5 = BruteForceIndex(Metric.from_string("Cosine"), FuturePool.immediate_pool(), device=local_rank, capacity_hint=0)
            g.manual_seed(0x5EED0007)
            for c0 in range(0, n5, 1_000_000):
                m = min(1_000_000, n5 - c0)
                rows = torch.randn((m, d), generator=g, device=dev) / d ** 0.5
                s_, e_ = max(c0, lo5), min(c0 + m, hi5)
                if s_ < e_:
                    ix5.append_batch_device(torch.arange(s_, e_, device=dev, dtype=torch.int64), rows[s_ - c0:e_ - c0].contiguous())
                del rows
            sx5 = ShardedBruteForceIndex(ix5, device=dev, share_seeds=not a.no_share_seeds, two_round=not a.one_round, push=not a.pull_bounds,
                                sliced_seeds=not a.full_seeds) if world > 1 else None
            rng5 = np.random.default_rng(0x5EED0008)            # the appended rows: same stream on every rank
            q5_pin = torch.empty((qb5, d), dtype=torch.float32).pin_memory()
            o5 = (torch.empty((qb5, k), dtype=torch.int64, device=dev), torch.empty((qb5, k), dtype=torch.float32, device=dev),
                  torch.empty((qb5,), dtype=torch.int32, device=dev))
            s0_, s1_ = (0, qb5) if world == 1 else sx5.slice_range(qb5)
            t_app = t_qry = 0.0
            appended = queried = 0
            visible = True
            n_b5 = (total5 + ab5 - 1) // ab5
            for i in range(n_b5 + 2):                               # two warm-up rounds, not timed
                timed_ = i >= 2
                m = min(ab5, total5 - appended) if timed_ else ab5
                if m <= 0:
                    break
                new_rows = (rng5.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
                base = n5 + 10 * total5 + i * ab5 if not timed_ else n5 + appended
                new_ids = np.arange(base, base + m, dtype=np.int64)
                barrier()
                t0 = time.perf_counter()
                if world == 1:
                    ix5.append_batch(new_ids, new_rows)             # host rows in: H2D + K1 inside the call
                else:
                    sx5.append_routed(new_ids, new_rows)            # round-robin by batch (ShardApi.scala:21-48)
                barrier()                                           # the JOB's clock: the batch is in place on whichever rank took it
                t1 = time.perf_counter()
                q5_pin.copy_(torch.from_numpy(new_rows[:qb5]))      # the queries ARE rows just appended: visibility check
                qd5 = q5_pin.to(dev, non_blocking=True)
                if world == 1:
                    ix5.query_batch_device(qd5, k, o5[0], o5[1], o5[2], stream.cuda_stream)
                    top1 = o5[0][:, 0].cpu()
                else:
                    top1 = sx5.batch_query_device(qd5, k, stream.cuda_stream, deliver="slice")[0][:, 0].cpu()
                t2 = time.perf_counter()
                visible &= bool((top1.numpy() == new_ids[s0_:s1_]).all())
                if timed_:
                    t_app += t1 - t0
                    t_qry += t2 - t1
                    appended += m
                    queried += qb5
            ix5.raise_pending_error()
            t_app, t_qry = max_over_ranks(t_app), max_over_ranks(t_qry)
            visible = max_over_ranks(0.0 if visible else 1.0) == 0.0
            extra["config5"] = {"workload": f"configs[4]: {n5}x{d} Cosine pre-loaded over {world} rank(s) into indexes created with capacity_hint = 0, "
                                            f"{appended} rows appended in host batches of {ab5}, a {qb5}-query batch after every append batch, top-{k}",
                                "n_gpus": world, "append_rows_per_s": appended / t_app, "queries_per_s": queried / t_qry,
                                "ms_per_query_batch": 1e3 * t_qry / max(1, queried // qb5), "ms_per_append_batch": 1e3 * t_app / max(1, n_b5),
                                "appended_rows_visible_to_next_query": visible, "mapped_bytes_rank0": ix5.stat("mapped_bytes"),
                                "algorithmic_bytes_rank0": ix5.stat("row_bytes") + ix5.stat("shadow_bytes"),
                                "timing": "host wall clock of the whole job: every append batch is timed to the barrier behind it (one rank takes it, the "
                                          "others wait), every query batch to the D2H of the answers; max over ranks"}
            ix5.close()
        except Exception as e:
            extra["config5"] = {"error": repr(e)}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only; bounded sample) + parity of the sample ----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            import oracle
            from oracle import METRIC_BY_NAME
            cores = oracle.max_threads()
            nq = a.cpu_queries or cores
            parts = []
            # the device-resident corpus itself, regenerated from the same stream: the CPU arm sees exactly the rows the GPU scanned
            g.manual_seed(0x5EED0001)
            for c0 in range(0, n, 1_000_000):
                m = min(1_000_000, n - c0)
                parts.append((torch.randn((m, d), generator=g, device=dev) / d ** 0.5).cpu())
            corpus_np = torch.cat(parts).numpy()
            del parts
            mo = METRIC_BY_NAME[a.metric]
            r = cpu_baseline_run(corpus_np, np.arange(n, dtype=np.int64), q_np[:nq], mo, k, 1, 0, keep=True)
            cpu = {kk: r[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            # parity of the sample: the CPU arm's own answers for these queries against the GPU batch's rows (outside every
            # timed region).  Two oracles: canonical (distance, id) order -- must match bit for bit -- and the reference-
            # faithful heap order, which may permute ids only inside runs of exactly equal distances.
            got = (full[0][:nq], full[1][:nq], full[2][:nq])
            canon = oracle.query_canonical(mo, corpus_np, None, q_np[:nq], k, nthreads=cores)
            d_ok, i_ok, _ = parity_against(got, canon)
            fd_ok, fi_ok, ties = parity_against(got, r["answers"], tie_tolerant=True)
            parity = {"queries": nq, "against": "oracle/oracle.c on the full corpus: canonical (C5) order and the reference-faithful "
                                                "linked-list + PriorityQueue restatement (the timed cpu_baseline run itself)",
                      "ids_identical": i_ok, "dist_bits_identical": d_ok, "faithful_dist_bits_identical": fd_ok,
                      "faithful_ids_identical_off_exact_ties": fi_ok, "faithful_tie_positions": ties}
            # separately labelled "fair CPU" figure: query-blocked, row-parallel, SIMD, -O3 -ffast-math (oracle/fast_cpu.c) --
            # so that the ratio is not only against the reference's pointer-chasing layout.  Timed only, never a parity oracle.
            nq2 = min(a.batch, 256)
            t0 = time.time()
            oracle.query_blocked_cpu(mo, corpus_np, None, q_np[:nq2], k, nthreads=cores)
            cpu["fair_blocked_simd_fp32"] = {"value": nq2 / (time.time() - t0), "unit": UNIT, "cores": cores, "queries": nq2,
                                             "what": "oracle/fast_cpu.c: 16-query blocks x 64-row tiles, threads split the rows, "
                                                     "gcc -O3 -ffast-math with AVX-512/AVX2 clones"}
            del corpus_np
        except Exception as e:  # the baseline is a report, never a reason to lose the measurement
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e!r}"}

    rc = 0
    if rank == 0:
        line = {
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if last_path == 2 else "f32",
            "dtype_detail": ("bf16 candidate filter on tcgen05 (fp32 accumulate)" if last_path == 2 else "f32 streaming scan")
                            + " + exact rescoring of the survivors: fp64 accumulation, one rounding to fp32 (bit-identical to the oracle)",
            "data": "synthetic", "config": config_dict(a, n_gpus, sx.route if sx else ""),
            "e2e": {"value": b / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": b * d * 4 * (world if (a.full_h2d or (sx and sx.route != "fused")) else 1),
                    "d2h_bytes_per_step": b * k * 12 + b * 4, "ms_per_step": e2e_ms,
                    "api": "ann_query_batch (host buffers)" if world == 1 else
                           ("per rank: pinned H2D of the whole batch" if a.full_h2d else "per rank: pinned H2D of the rank's 1/N slice of the batch + NVLink push all-gather")
                           + " + three-phase sharded query + slice merge + D2H of the rank's 1/N slice of the answer"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "path": {1: "scan", 2: "gemm"}.get(last_path, str(last_path)),
            "result_digest": digest, "flagged_rows": flagged_rows, "parity_sample": parity, "breakdown": breakdown, "extra": extra,
        }
        print(json.dumps(line), flush=True)
        if flagged_rows:
            print("bench.py: the timed batch contains flagged (invalid) rows", file=sys.stderr)
            rc = 3
        if parity and not (parity["ids_identical"] and parity["dist_bits_identical"] and parity["faithful_dist_bits_identical"]
                           and parity["faithful_ids_identical_off_exact_ties"]):
            print("bench.py: PARITY MISMATCH between the GPU batch and the CPU oracle sample", file=sys.stderr)
            rc = 4
    if not ix_closed:
        ix.close()
    if world > 1:
        dist.destroy_process_group()
    return rc


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
