"""Multi-GPU parity check, run under torchrun on R GPUs of one box:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Rows are sharded contiguously across ranks, queries replicated, each rank answers locally through the C ABI
(ann_query_batch_device), the per-rank top-k are all-gathered over NCCL and merged by the K5 kernel
(ann_merge_topk_device).  Every rank must hold the single-shard answer of the CPU oracle, bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
import oracle  # noqa: E402
from the_algorithm_b200.ann.brute_force import BruteForceIndex, merge_topk_device  # noqa: E402
from the_algorithm_b200.ann.common import Cosine, FuturePool, InnerProduct, L2  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok_all = True
for metric, n, d, b, k in ((InnerProduct, 200_003, 200, 300, 100), (Cosine, 50_000, 64, 40, 10), (L2, 120_000, 128, 5, 100)):
    rng = np.random.default_rng(7)
    corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    corpus[n // 2: n // 2 + 50] = corpus[:50]            # exact ties across shard boundaries
    ids = rng.permutation(n).astype(np.int64)
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    lo, hi = rank * n // world, (rank + 1) * n // world
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool(), device=local)
    ix.append_batch(ids[lo:hi], corpus[lo:hi])
    qd = torch.from_numpy(q).to(dev)
    oi = torch.empty((b, k), dtype=torch.int64, device=dev)
    od = torch.empty((b, k), dtype=torch.float32, device=dev)
    oc = torch.empty((b,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ix.query_batch_device(qd, k, oi, od, oc, st)
    g_ids = torch.empty((world, b, k), dtype=torch.int64, device=dev)
    g_dist = torch.empty((world, b, k), dtype=torch.float32, device=dev)
    g_cnt = torch.empty((world, b), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(g_ids, oi)
    dist.all_gather_into_tensor(g_dist, od)
    dist.all_gather_into_tensor(g_cnt, oc)
    mi, md, mc = merge_topk_device(g_ids, g_dist, g_cnt, k, st)
    torch.cuda.synchronize()
    ix.raise_pending_error()
    wi, wd, wc = oracle.query_canonical(metric.ordinal, corpus, ids, q, k)
    ok = bool((mi.cpu().numpy() == wi).all() and (md.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
              and (mc.cpu().numpy() == wc).all())
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{metric.name} n={n} d={d} b={b} k={k} world={world}: identical_to_single_shard_oracle={bool(t.item())}", flush=True)
    ok_all &= bool(t.item())
    ix.close()
dist.destroy_process_group()
if rank == 0:
    print("DIST_OK", ok_all, flush=True)
sys.exit(0 if ok_all else 1)
