"""Developer check run on a B200 through gpurun: parity of the streaming-scan path against the CPU oracle on a
spread of shapes, then raw timing at the BASELINE config-3 shape (10M x 128, L2, batch 1).  Writes
gpurun_out/dev_check_scan.json.  Not part of the test suite."""
from __future__ import annotations

import json
import os
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

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
res = {"parity": [], "timing": []}


def parity(metric, n, d, b, k, seed, dup=False, ids_mode="iota"):
    rng = np.random.default_rng(seed)
    corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    if dup and n > 10:
        corpus[n // 2:n // 2 + n // 100 + 1] = corpus[: n // 100 + 1]
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.int64) if ids_mode == "iota" else rng.permutation(n).astype(np.int64) * 7 - 3
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool())
    t0 = time.time()
    ix.append_batch(ids, corpus)
    gi, gd, gc = ix.batch_query_with_distance(q, k)
    t1 = time.time()
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, k)
    ok_ids = bool((gi == oi).all())
    ok_dist = bool((gd.view(np.uint32) == od.view(np.uint32)).all())
    ok_cnt = bool((gc == oc).all())
    rec = dict(metric=metric.name, n=n, d=d, b=b, k=k, dup=dup, ids=ids_mode, ids_equal=ok_ids, dist_bits_equal=ok_dist,
               count_equal=ok_cnt, gpu_s=round(t1 - t0, 4), launches=ix.stat("launches"))
    if not ok_ids:
        bad = np.argwhere(gi != oi)
        rec["first_bad"] = bad[:3].tolist()
    res["parity"].append(rec)
    print(rec, flush=True)
    ix.close()
    return ok_ids and ok_dist and ok_cnt


all_ok = True
for metric in ((InnerProduct, Cosine, L2) if "--no-parity" not in sys.argv else ()):
    for (n, d, b, k) in [(1000, 16, 3, 10), (5000, 200, 9, 100), (100_000, 200, 16, 100), (70_001, 128, 5, 100),
                         (3000, 100, 4, 7), (257, 36, 2, 300), (50, 8, 1, 100), (200_000, 64, 1, 200)]:
        all_ok &= parity(metric, n, d, b, k, seed=n + d)
    all_ok &= parity(metric, 20_000, 200, 8, 100, seed=5, dup=True, ids_mode="perm")
res["all_ok"] = bool(all_ok)

# ---- timing at config 3 (10M x 128 L2, batch 1) and at 10M x 200 IP -----------------------------------------------
import torch  # noqa: E402

for (metric, n, d) in [(L2, 10_000_000, 128), (InnerProduct, 10_000_000, 200)]:
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    ix = BruteForceIndex(metric, FuturePool.immediate_pool(), capacity_hint=n)
    chunk = 1_000_000
    for c0 in range(0, n, chunk):
        rows = torch.randn((chunk, d), generator=g, device=dev, dtype=torch.float32) / (d ** 0.5)
        ids = torch.arange(c0, c0 + chunk, device=dev, dtype=torch.int64)
        ix.append_batch_device(ids, rows)
    del rows
    for b in (1, 2, 4, 8):
        q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).contiguous()
        oi = torch.empty((b, 100), dtype=torch.int64, device=dev)
        od = torch.empty((b, 100), dtype=torch.float32, device=dev)
        oc = torch.empty((b,), dtype=torch.int32, device=dev)
        ts = torch.cuda.current_stream()
        st = ts.cuda_stream
        torch.cuda.synchronize()
        for _ in range(3):
            ix.query_batch_device(q, 100, oi, od, oc, st)
        torch.cuda.synchronize()
        ix.raise_pending_error()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(reps):
            ix.query_batch_device(q, 100, oi, od, oc, st)
        e1.record(ts)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = n * d * 4 / (ms * 1e-3) / 1e9
        rec = dict(metric=metric.name, n=n, d=d, b=b, ms_per_batch=round(ms, 4), scan_GBps=round(gbs, 1),
                   qps=round(b / (ms * 1e-3), 1))
        res["timing"].append(rec)
        print(rec, flush=True)
    # spot parity at full size against the fast CPU scan on 2 queries (ids only; fp32 CPU path is not bit-exact)
    qh = q[:2].cpu().numpy()
    gi = oi[:2].cpu().numpy()
    ix.close()
    del ix
    torch.cuda.empty_cache()

(OUT / "dev_check_scan.json").write_text(json.dumps(res, indent=1))
print("ALL_OK", all_ok)
sys.exit(0 if all_ok else 1)
