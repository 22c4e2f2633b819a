"""The ONE-PROCESS sharded handle (`ann_sharded_*`, the form a single-JVM host binds) on the real GPUs of a box:
parity against a single index on GPU 0 and end-to-end throughput through `ann_sharded_query_batch` (host buffers in,
host buffers out).  Run under gpurun --gpus R:

    python tools/sharded_handle_bench.py [rows] [batch] [dim] [steps] [devices, e.g. 0,1 or 0,0 for two shards on GPU 0] [pageable,pinned]
"""
from __future__ import annotations

import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
import torch  # noqa: E402

from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import FuturePool, InnerProduct  # noqa: E402
from the_algorithm_b200.ann.sharded import GpuShardedBruteForceIndex  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 200
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
k = 100
devs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else list(range(torch.cuda.device_count()))
R = len(devs)
rng = np.random.default_rng(5)


def digest(ids, dist, cnt) -> str:
    h = hashlib.sha256()
    for a in (ids, dist.view(np.uint32), cnt):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:32]


one = BruteForceIndex(InnerProduct, FuturePool.immediate_pool(), device=0, capacity_hint=n)
sx = GpuShardedBruteForceIndex(InnerProduct, FuturePool.immediate_pool(), dim=d, devices=devs, capacity_hint=n)
for c0 in range(0, n, 500_000):
    m = min(500_000, n - c0)
    rows = (rng.standard_normal((m, d), dtype=np.float32) / np.float32(np.sqrt(d)))
    ids = np.arange(c0, c0 + m, dtype=np.int64)
    one.append_batch(ids, rows)
    sx.append_batch(ids, rows)
q0 = rng.uniform(-1, 1, (b, d)).astype(np.float32)
want = one.batch_query_with_distance(q0, k)
modes = sys.argv[6].split(",") if len(sys.argv) > 6 else ["pageable", "pinned"]   # caller buffers: pageable (a JVM direct buffer) / page-locked
ok = True
for mode in modes:
    pinned = mode == "pinned"
    q = torch.from_numpy(q0).pin_memory().numpy() if pinned else q0
    mk = (lambda *s_, dtype: torch.empty(s_, dtype=dtype).pin_memory().numpy()) if pinned else (lambda *s_, dtype: torch.empty(s_, dtype=dtype).numpy())
    outs = (mk(b, k, dtype=torch.int64), mk(b, k, dtype=torch.float32), mk(b, dtype=torch.int32))
    out = {"caller_buffers": mode, "rows": n, "batch": b, "dim": d, "k": k, "devices": devs, "shard_sizes": sx.shard_sizes(),
           "peer_access": sx.stat("peer_access")}
    for sliced in (1, 0):
        sx.set_option("sliced_seeds", sliced)
        got = sx.batch_query_with_distance(q, k)
        same = bool((got[0] == want[0]).all() and (got[1].view(np.uint32) == want[1].view(np.uint32)).all() and (got[2] == want[2]).all())
        for _ in range(3):
            sx.batch_query_with_distance(q, k, out=outs)
        t0 = time.perf_counter()
        for _ in range(steps):
            got = sx.batch_query_with_distance(q, k, out=outs)
        dt = (time.perf_counter() - t0) / steps
        same = same and bool((got[0] == want[0]).all())
        out["sliced_seeds" if sliced else "full_seeds"] = {"identical_to_single_index": same, "ms_per_batch_host_to_host": dt * 1e3,
                                                           "queries_per_s": b / dt, "digest": digest(*got)}
        ok = ok and same
    for _ in range(3):
        one.batch_query_with_distance(q, k, out=outs)
    t0 = time.perf_counter()
    for _ in range(steps):
        one.batch_query_with_distance(q, k, out=outs)
    dt1 = (time.perf_counter() - t0) / steps
    out["single_index_gpu0"] = {"ms_per_batch_host_to_host": dt1 * 1e3, "queries_per_s": b / dt1, "digest": digest(*want)}
    out["speedup_over_single_index"] = dt1 / (out["sliced_seeds"]["ms_per_batch_host_to_host"] * 1e-3)
    out["fallback_batches"] = sx.stat("fallback_batches")
    print(json.dumps(out), flush=True)
print("SHARDED_HANDLE_OK" if ok else "SHARDED_HANDLE_MISMATCH")
sys.exit(0 if ok else 1)
