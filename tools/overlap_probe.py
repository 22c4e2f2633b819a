"""Probe (run under gpurun): do two query batches in flight on two streams overlap one batch's selection kernels
(compaction, finalize) with the other batch's tensor-core filter?

    python tools/overlap_probe.py [n] [b] [d] [reps]

Two handles over the same rows stand in for two scratch sets of one handle.  Prints ms per batch for one stream
and for two alternating streams, same process, alternating visits.
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
from the_algorithm_b200.ann.common import FuturePool, InnerProduct  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 200
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
k = 100
dev = torch.device("cuda", 0)


def build():
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    ix = BruteForceIndex(InnerProduct, FuturePool.immediate_pool(), capacity_hint=n)
    for c0 in range(0, n, 1_000_000):
        m = min(1_000_000, n - c0)
        rows = torch.randn((m, d), generator=g, device=dev, dtype=torch.float32) / (d ** 0.5)
        ix.append_batch_device(torch.arange(c0, c0 + m, device=dev, dtype=torch.int64), rows)
    ix.set_option("path", 2)
    return ix


ixs = [build(), build()]
g = torch.Generator(device=dev)
g.manual_seed(2)
q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
outs = [(torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev),
         torch.empty((b,), dtype=torch.int32, device=dev)) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for i in range(2):
    for _ in range(3):
        ixs[i].query_batch_device(q, k, *outs[i], streams[i].cuda_stream)
torch.cuda.synchronize()
assert torch.equal(outs[0][0], outs[1][0])


def run(two: bool) -> float:
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    streams[1].wait_event(e0)
    for r in range(reps):
        i = (r & 1) if two else 0
        ixs[i].query_batch_device(q, k, *outs[i], streams[i].cuda_stream)
    streams[0].wait_stream(streams[1])
    e1.record(streams[0])
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t = {False: [], True: []}
for _ in range(5):
    for two in (False, True):
        t[two].append(run(two))
for two in (False, True):
    print(f"n={n} b={b} streams={'2' if two else '1'}: median {statistics.median(t[two]):.3f} ms/batch  visits " + " ".join(f"{x:.3f}" for x in t[two]), flush=True)
for ix in ixs:
    ix.raise_pending_error()
print("PROBE_OK")
