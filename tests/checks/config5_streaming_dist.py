"""BASELINE.json config 5 on R GPUs of one box (run under torchrun): streaming Appendable interleaved with batched queries.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 \
        tests/checks/config5_streaming_dist.py [--rows 50000000 --appends 1000000 --append-batch 4096 --query-batch 256]

A `rows` x 200 index is pre-loaded row-sharded over the ranks; then `appends` rows arrive in host batches, routed to the
shards round-robin by batch (ShardedAppendable with the deterministic stand-in for RandomShardFunction, ShardApi.scala:21-48),
and after every append batch a batch of `query-batch` queries is answered by all ranks (two-phase shard query with shared
seed thresholds + fused exchange/merge).  The queries are the first rows of the batch just appended, so the visibility
contract is checked on every rank for every batch: a row whose append returned before the query was issued is found by it
(BruteForceIndex.scala:34-36 gives the same guarantee), here as its own nearest neighbour under Cosine.
Reports (max over ranks): appended rows/s, queries/s, ms per query batch, and the visibility verdict.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import FuturePool, Metric  # noqa: E402
from the_algorithm_b200.ann.distributed import ShardedBruteForceIndex, shard_range  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--dim", type=int, default=200)
ap.add_argument("--metric", default="Cosine")
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--appends", type=int, default=1_000_000)
ap.add_argument("--append-batch", type=int, default=4096)
ap.add_argument("--query-batch", type=int, default=256)
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
metric = Metric.from_string(a.metric)
n, d, k, ab, qb = a.rows, a.dim, a.k, a.append_batch, a.query_batch
lo, hi = shard_range(n, world, rank)
ix = BruteForceIndex(metric, FuturePool.immediate_pool(), device=local, capacity_hint=(hi - lo) + a.appends // world + 2 * ab)
g = torch.Generator(device=dev)
g.manual_seed(0x5EED0001)
for c0 in range(0, n, 1_000_000):      # same generator stream on every rank, each keeps its own row range (as bench.py)
    m = min(1_000_000, n - c0)
    rows = torch.randn((m, d), generator=g, device=dev) / d ** 0.5
    s, e = max(c0, lo), min(c0 + m, hi)
    if s < e:
        ix.append_batch_device(torch.arange(s, e, device=dev, dtype=torch.int64), rows[s - c0:e - c0].contiguous())
del rows
sx = ShardedBruteForceIndex(ix, device=dev)
stream = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0x5EED0005)            # the appended rows: same stream on every rank

n_batches = (a.appends + ab - 1) // ab
visible = True
t_append = t_query = 0.0
appended = queried = 0
q_pin = torch.empty((qb, d), dtype=torch.float32).pin_memory()
for i in range(n_batches + 2):                      # two warm-up rounds, not timed
    timed = i >= 2
    m = min(ab, a.appends - appended) if timed else ab
    if m <= 0:
        break
    new_rows = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    base = n + 10 * a.appends + i * ab if not timed else n + appended
    new_ids = np.arange(base, base + m, dtype=np.int64)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sx.append_routed(new_ids, new_rows)             # host rows -> the shard this batch is routed to (H2D + K1 inside)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    nq = min(qb, m)
    q_pin[:nq].copy_(torch.from_numpy(new_rows[:nq]))
    q_dev = q_pin.to(dev, non_blocking=True) if nq == qb else q_pin[:nq].to(dev)
    oi, od, oc = sx.batch_query_device(q_dev, k, stream)
    top1 = oi[:, 0].cpu()                           # D2H + synchronise: the answer is on the host
    t2 = time.perf_counter()
    visible &= bool((top1.numpy() == new_ids[:nq]).all())
    if timed:
        t_append += t1 - t0
        t_query += t2 - t1
        appended += m
        queried += nq
ix.raise_pending_error()
t = torch.tensor([t_append, t_query, 0.0 if visible else 1.0], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
size = torch.tensor([ix.size()], device=dev, dtype=torch.int64)
dist.all_reduce(size, op=dist.ReduceOp.SUM)
if rank == 0:
    rec = {"config": f"5: {n}x{d} {a.metric} pre-loaded over {world} GPUs + {appended} rows appended in host batches of {ab} "
                     f"(round-robin by batch), a {qb}-query batch after every append batch, top-{k}",
           "n_gpus": world, "append_rows_per_s": appended / float(t[0]), "queries_per_s": queried / float(t[1]),
           "ms_per_query_batch": 1e3 * float(t[1]) / max(1, queried // qb), "ms_per_append_batch": 1e3 * float(t[0]) / n_batches,
           "appended_rows_visible_to_next_query": bool(t[2].item() == 0.0), "final_rows": int(size.item()),
           "route": sx.route, "shared_seed_thresholds": sx.share_seeds,
           "timing": "host clock around each call incl. H2D of the appended rows / D2H of the answers, max over ranks"}
    print(json.dumps(rec), flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"config5_streaming_n{world}.json").write_text(json.dumps(rec) + "\n")
ix.close()
dist.destroy_process_group()
sys.exit(0 if bool(t[2].item() == 0.0) else 1)
