"""One rank's share of a row-sharded query, with and without shared seed thresholds, measured on ONE GPU.

    python tools/seed_share_bench.py [world=8] [rows_total=10000000] [b=4096] [d=200] [reps=10] [metric=InnerProduct]

`world` shard indices of rows_total/world rows each are built on one device (10M x 200 in 8 shards = the per-rank state of
`bench.py --gpus 8`).  Timed with CUDA events on the launching stream:
  plain   : ann_query_batch_device on shard 0                      (what every rank did before)
  shared  : ann_query_seed_device + ann_query_finish_device on shard 0, against the bounds all `world` shards published
            (the other shards' seed calls are not in the timed region: on a real box they run on the other GPUs)
and, for the whole emulated box, that the merged two-phase answer equals the merged plain answer bit for bit.
"""
from __future__ import annotations

import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
import torch  # noqa: E402

from the_algorithm_b200.ann.brute_force import BruteForceIndex, merge_topk_device  # noqa: E402
from the_algorithm_b200.ann.common import FuturePool, Metric  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
b = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
d = int(sys.argv[4]) if len(sys.argv) > 4 else 200
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
metric = Metric.from_string(sys.argv[6]) if len(sys.argv) > 6 else Metric.from_string("InnerProduct")
k = 100

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator(device=dev)
g.manual_seed(0x5EED0001)
n_per = n // world
shards = []
for s in range(world):
    ix = BruteForceIndex(metric, FuturePool.immediate_pool(), device=0, capacity_hint=n_per)
    for c0 in range(0, n_per, 1_000_000):
        m = min(1_000_000, n_per - c0)
        rows = torch.randn((m, d), generator=g, device=dev) / d ** 0.5
        ix.append_batch_device(torch.arange(s * n_per + c0, s * n_per + c0 + m, device=dev, dtype=torch.int64), rows)
    shards.append(ix)
g.manual_seed(0x5EED0002)
q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
stream = torch.cuda.current_stream()
st = stream.cuda_stream

keys = torch.empty((world, b, k), dtype=torch.int32, device=dev)
ptrs = [keys[s].data_ptr() for s in range(world)]
res = [torch.empty((world, b, k), dtype=torch.int64, device=dev), torch.empty((world, b, k), dtype=torch.float32, device=dev),
       torch.empty((world, b), dtype=torch.int32, device=dev)]


def plain(s):
    shards[s].query_batch_device(q, k, res[0][s], res[1][s], res[2][s], st)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- whole emulated box once, both ways: identical merged answers ----
for s in range(world):
    plain(s)
want = [t.clone() for t in merge_topk_device(res[0], res[1], res[2], k)]
for s in range(world):
    shards[s].query_seed_device(q, k, keys[s], st)
for s in range(world):
    shards[s].query_finish_device(q, k, ptrs, res[0][s], res[1][s], res[2][s], st)
got = merge_topk_device(res[0], res[1], res[2], k)
torch.cuda.synchronize()
for ix in shards:
    ix.raise_pending_error()
same = all(bool((a == c).all()) for a, c in zip(got, want))
mean_cnt = float(res[2].float().mean())
print(f"world={world} rows/shard={n_per} b={b} d={d} {metric}: merged answers identical: {same}; "
      f"mean entries a shard contributes per query with sharing: {mean_cnt:.1f} of {k}")
assert same

# ---- shard 0, alternating, same box / clocks ----
tp, ts = [], []
for r in range(5):
    tp.append(timed(lambda: plain(0)))
    chunks_plain = shards[0].stat("last_gemm_chunks")

    def shared():
        shards[0].query_seed_device(q, k, keys[0], st)
        shards[0].query_finish_device(q, k, ptrs, res[0][0], res[1][0], res[2][0], st)

    ts.append(timed(shared))
    chunks_shared = shards[0].stat("last_gemm_chunks")
print("plain  ms/batch:", " ".join(f"{x:.3f}" for x in tp), "median", f"{statistics.median(tp):.3f}", "chunks", chunks_plain)
print("shared ms/batch:", " ".join(f"{x:.3f}" for x in ts), "median", f"{statistics.median(ts):.3f}", "chunks", chunks_shared)
print(f"per-rank speed-up from shared seeds: {statistics.median(tp) / statistics.median(ts):.3f}x")
for ix in shards:
    ix.close()
