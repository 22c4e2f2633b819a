"""Measures both query paths over a grid of (rows, batch) on one B200 and prints a table (run under gpurun)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import FuturePool, InnerProduct  # noqa: E402

dev = torch.device("cuda", 0)
d, k = 200, 100
print(f"# rows x {d} fp32, InnerProduct top-{k}; ms per call (device-resident, CUDA events), scan path vs tensor-core path")
print(f"{'rows':>10} {'batch':>6} {'scan_ms':>9} {'gemm_ms':>9}")
for n in (100_000, 1_000_000, 10_000_000):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    ix = BruteForceIndex(InnerProduct, FuturePool.immediate_pool(), capacity_hint=n)
    for c0 in range(0, n, 1_000_000):
        m = min(1_000_000, n - c0)
        ix.append_batch_device(torch.arange(c0, c0 + m, device=dev, dtype=torch.int64),
                               torch.randn((m, d), generator=g, device=dev) / d ** 0.5)
    for b in (1, 2, 4, 8, 16, 32, 64, 256):
        q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
        oi = torch.empty((b, k), dtype=torch.int64, device=dev)
        od = torch.empty((b, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream()
        res = {}
        for path in (1, 2):
            if path == 1 and b > 64:
                res[path] = float("nan")
                continue
            ix.set_option("path", path)
            for _ in range(3):
                ix.query_batch_device(q, k, oi, od, None, st.cuda_stream)
            torch.cuda.synchronize()
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                ix.query_batch_device(q, k, oi, od, None, st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            ix.raise_pending_error()
            res[path] = e0.elapsed_time(e1) / reps
        print(f"{n:>10} {b:>6} {res[1]:>9.3f} {res[2]:>9.3f}", flush=True)
    ix.close()
    del ix
    torch.cuda.empty_cache()
