"""Developer check for the tcgen05 GEMM-filter path (run under gpurun, each invocation wrapped in `timeout`).

    python tests/checks/dev_check_gemm.py parity <cta_group>        small/medium parity against the CPU oracle
    python tests/checks/dev_check_gemm.py time <cta_group> [n] [b]  timing at the headline shape (10M x 200, b = 4096)
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
import oracle  # noqa: E402
from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import Cosine, FuturePool, InnerProduct, L2  # noqa: E402

mode = sys.argv[1]
cg = int(sys.argv[2])


def parity(metric, n, d, b, k, seed, dup=False):
    rng = np.random.default_rng(seed)
    corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    if dup:
        corpus[n // 2:n // 2 + n // 100 + 1] = corpus[: n // 100 + 1]
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    ids = rng.permutation(n).astype(np.int64) * 5 - 11
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool())
    ix.append_batch(ids, corpus)
    ix.set_option("path", 2)
    ix.set_option("gemm_cta_group", cg)
    t0 = time.time()
    gi, gd, gc = ix.batch_query_with_distance(q, k)
    t1 = time.time()
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, k)
    ok = bool((gi == oi).all() and (gd.view(np.uint32) == od.view(np.uint32)).all() and (gc == oc).all())
    rec = dict(metric=metric.name, n=n, d=d, b=b, k=k, cg=cg, ok=ok, ids_equal=bool((gi == oi).all()),
               gpu_s=round(t1 - t0, 4), path=ix.stat("last_path"), launches=ix.stat("launches"))
    if not ok:
        bad = np.argwhere(gi != oi)
        rec["n_bad"] = int(bad.shape[0])
        rec["first_bad"] = bad[:3].tolist()
        if bad.shape[0]:
            qq, jj = bad[0]
            rec["gpu_row"] = gi[qq, max(0, jj - 1):jj + 3].tolist()
            rec["ora_row"] = oi[qq, max(0, jj - 1):jj + 3].tolist()
            rec["gpu_d"] = gd[qq, max(0, jj - 1):jj + 3].tolist()
            rec["ora_d"] = od[qq, max(0, jj - 1):jj + 3].tolist()
    print(rec, flush=True)
    ix.close()
    return ok


if mode == "parity":
    ok = True
    ok &= parity(InnerProduct, 4096, 64, 128, 10, 1)
    ok &= parity(InnerProduct, 5000, 128, 130, 100, 2)
    for metric in (InnerProduct, Cosine, L2):
        ok &= parity(metric, 50_000, 200, 300, 100, 3)
        ok &= parity(metric, 33_333, 128, 257, 100, 4, dup=True)
        ok &= parity(metric, 20_000, 72, 64, 17, 5)
    print("PARITY_OK", ok)
    sys.exit(0 if ok else 1)

if mode == "time":
    import torch

    n = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
    b = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    d, k = (int(sys.argv[5]) if len(sys.argv) > 5 else 200), 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    ix = BruteForceIndex(InnerProduct, FuturePool.immediate_pool(), capacity_hint=n)
    chunk = 1_000_000
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        rows = torch.randn((m, d), generator=g, device=dev, dtype=torch.float32) / (d ** 0.5)
        ix.append_batch_device(torch.arange(c0, c0 + m, device=dev, dtype=torch.int64), rows)
    del rows
    ix.set_option("path", 2)
    ix.set_option("gemm_cta_group", cg)
    q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
    oi = torch.empty((b, k), dtype=torch.int64, device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    oc = torch.empty((b,), dtype=torch.int32, device=dev)
    ts = torch.cuda.current_stream()
    st = ts.cuda_stream
    torch.cuda.synchronize()
    for _ in range(2):
        ix.query_batch_device(q, k, oi, od, oc, st)
    torch.cuda.synchronize()
    ix.raise_pending_error()
    reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
    for _ in range(reps // 2):   # thermal / power steady state before timing long runs
        ix.query_batch_device(q, k, oi, od, oc, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(reps):
        ix.query_batch_device(q, k, oi, od, oc, st)
    e1.record(ts)
    torch.cuda.synchronize()
    ix.raise_pending_error()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * n * d * b / (ms * 1e-3) / 1e12
    print(dict(n=n, d=d, b=b, cg=cg, ms_per_batch=round(ms, 3), qps=round(b / (ms * 1e-3)), tflops=round(tf, 1)), flush=True)
    # cross-check a few queries against the scan path (itself oracle-verified)
    gi = oi[:8].cpu().numpy()
    gd = od[:8].cpu().numpy()
    ix.set_option("path", 1)
    oi2 = torch.empty((8, k), dtype=torch.int64, device=dev)
    od2 = torch.empty((8, k), dtype=torch.float32, device=dev)
    ix.query_batch_device(q[:8].contiguous(), k, oi2, od2, None, st)
    torch.cuda.synchronize()
    same = bool((oi2.cpu().numpy() == gi).all() and (od2.cpu().numpy().view(np.uint32) == gd.view(np.uint32)).all())
    print("GEMM_VS_SCAN_SAME", same, flush=True)
    sys.exit(0 if same else 1)
