"""Same-process A/B of one engine option on the GEMM-filter path (run under gpurun).

    python tools/ab_options.py OPTION v1,v2[,v3] [n] [b] [d] [rounds] [reps] [metric]

One index is built once; the option values are then timed in alternation (v1 v2 v3 v1 v2 v3 ...), `reps` batches per
visit with CUDA events on the launching stream, so every value sees the same box, clocks and thermal state.  The GPU is
power-capped, which makes numbers taken in different processes differ by a few percent: this is the only A/B that can
resolve 1 % effects.  Prints the per-visit times and the median per value.
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

from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import Cosine, FuturePool, InnerProduct, L2  # noqa: E402

opt = sys.argv[1]
values = [int(v) for v in sys.argv[2].split(",")]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
b = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
d = int(sys.argv[5]) if len(sys.argv) > 5 else 200
rounds = int(sys.argv[6]) if len(sys.argv) > 6 else 5
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
metric = {"ip": InnerProduct, "cosine": Cosine, "l2": L2}[sys.argv[8] if len(sys.argv) > 8 else "ip"]
k = 100

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
ix = BruteForceIndex(metric, FuturePool.immediate_pool(), capacity_hint=n)
for c0 in range(0, n, 1_000_000):
    m = min(1_000_000, n - c0)
    rows = torch.randn((m, d), generator=g, device=dev, dtype=torch.float32) / (d ** 0.5)
    ix.append_batch_device(torch.arange(c0, c0 + m, device=dev, dtype=torch.int64), rows)
del rows
ix.set_option("path", 2)
q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
oi = torch.empty((b, k), dtype=torch.int64, device=dev)
od = torch.empty((b, k), dtype=torch.float32, device=dev)
oc = torch.empty((b,), dtype=torch.int32, device=dev)
ts = torch.cuda.current_stream()
st = ts.cuda_stream
ref = None
times = {v: [] for v in values}
for v in values:   # warm every variant (module load, I-cache) and check they agree
    ix.set_option(opt, v)
    for _ in range(3):
        ix.query_batch_device(q, k, oi, od, oc, st)
    torch.cuda.synchronize()
    ix.raise_pending_error()
    cur = (oi.clone(), od.clone())
    if ref is None:
        ref = cur
    elif not (torch.equal(ref[0], cur[0]) and torch.equal(ref[1].view(torch.int32), cur[1].view(torch.int32))):
        print("RESULTS_DIFFER", opt, v)
        sys.exit(1)
for r in range(rounds):
    for v in values:
        ix.set_option(opt, v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(reps):
            ix.query_batch_device(q, k, oi, od, oc, st)
        e1.record(ts)
        torch.cuda.synchronize()
        times[v].append(e0.elapsed_time(e1) / reps)
ix.raise_pending_error()
for v in values:
    print(f"{opt}={v}: median {statistics.median(times[v]):.3f} ms  visits " + " ".join(f"{t:.3f}" for t in times[v]), flush=True)
print("AB_OK")
