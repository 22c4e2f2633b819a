"""Multi-GPU parity check, run under torchrun on R GPUs of one box:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/checks/dist_check.py

Rows are sharded contiguously across ranks, queries replicated, each rank answers locally through the C ABI
(ann_query_batch_device), the per-rank top-k are all-gathered over NCCL and merged by the K5 kernel
(ann_merge_topk_device) -- and, second route, exchanged and merged by the fused peer-memory kernel
(ann_exchange_merge_device, ann/exchange.py).  Every rank must hold the single-shard answer of the CPU oracle, bit for
bit, by both routes; the two routes are also timed (CUDA events, max over ranks)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
import oracle  # noqa: E402
from the_algorithm_b200.ann.brute_force import BruteForceIndex, merge_topk_device  # noqa: E402
from the_algorithm_b200.ann.common import Cosine, FuturePool, InnerProduct, L2  # noqa: E402
from the_algorithm_b200.ann.distributed import ShardedBruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.exchange import PeerExchange  # noqa: E402

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
    # second route: the shard's results land in a peer-mapped block; one fused kernel exchanges and merges
    px = PeerExchange(b, k, dev)
    ok2 = True
    for rep in range(3):   # repeated: the barriers must also protect block reuse
        ix.query_batch_device(qd, k, *px.local.tensors, st)
        fi, fd, fc = px.exchange_merge(st)
        torch.cuda.synchronize()
        ok2 &= bool((fi.cpu().numpy() == wi).all() and (fd.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
                    and (fc.cpu().numpy() == wc).all())
    ix.raise_pending_error()
    # third: the host class that bench.py uses (route picked automatically; must be the fused one on this box)
    sx = ShardedBruteForceIndex(ix, device=dev)
    for rep in range(2):
        si, sd, sc = sx.batch_query_device(qd, k, st)
        torch.cuda.synchronize()
        ok2 &= bool((si.cpu().numpy() == wi).all() and (sd.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
                    and (sc.cpu().numpy() == wc).all())
    ok2 &= sx.route == "fused" and sx.share_seeds and sx.two_round
    # ... and its slice delivery: every rank receives only its rows of the merged batch, nothing is pushed
    q0, q1 = sx.slice_range(b)
    for rep in range(3):
        si, sd, sc = sx.batch_query_device(qd, k, st, deliver="slice")
        torch.cuda.synchronize()
        ok2 &= bool(si.shape[0] == q1 - q0 and (si.cpu().numpy() == wi[q0:q1]).all()
                    and (sd.cpu().numpy().view(np.uint32) == wd[q0:q1].view(np.uint32)).all() and (sc.cpu().numpy() == wc[q0:q1]).all())
    # ... fed by the partitioned query batch: every rank copies only its slice from the host, the rest arrives over NVLink
    if d % 4 == 0:
        q_pin = torch.from_numpy(q).pin_memory()
        for rep in range(2):
            qg = sx.gather_queries(q_pin[q0:q1], b, k, st)
            si, sd, sc = sx.batch_query_device(qg, k, st, deliver="slice")
            torch.cuda.synchronize()
            ok2 &= bool(torch.equal(qg.cpu(), q_pin) and (si.cpu().numpy() == wi[q0:q1]).all()
                        and (sd.cpu().numpy().view(np.uint32) == wd[q0:q1].view(np.uint32)).all())
    del sx
    # the same three phases with PULL delivery of the bounds (consumers read the peers' arrays over NVLink)
    sx = ShardedBruteForceIndex(ix, device=dev, push=False)
    for rep in range(2):
        si, sd, sc = sx.batch_query_device(qd, k, st)
        torch.cuda.synchronize()
        ok2 &= bool((si.cpu().numpy() == wi).all() and (sd.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
                    and (sc.cpu().numpy() == wc).all())
    del sx
    # push delivery with every shard seeding every query (k bounds per query; A/B of sliced seeding, the default above)
    sx = ShardedBruteForceIndex(ix, device=dev, sliced_seeds=False)
    for rep in range(2):
        si, sd, sc = sx.batch_query_device(qd, k, st)
        torch.cuda.synchronize()
        ok2 &= bool((si.cpu().numpy() == wi).all() and (sd.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
                    and (sc.cpu().numpy() == wc).all())
    del sx
    # the same with the seed round only (A/B of the second cross-shard round)
    sx = ShardedBruteForceIndex(ix, device=dev, two_round=False)
    for rep in range(2):
        si, sd, sc = sx.batch_query_device(qd, k, st)
        torch.cuda.synchronize()
        ok2 &= bool((si.cpu().numpy() == wi).all() and (sd.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
                    and (sc.cpu().numpy() == wc).all())
    del sx
    # degenerate batch: a NaN query and a query nearest to thousands of identical rows flag some shard's bounded selector.
    # The asynchronous call must say so (count = -1 on every rank), the synchronous one must answer exactly.
    if metric is InnerProduct:
        corpus2 = corpus.copy()
        corpus2[lo + 100: lo + 100 + min(7000, hi - lo - 100)] = corpus2[3]      # every shard holds thousands of duplicates
        q2 = q.copy()
        q2[1] = corpus2[3] * 4.0
        q2[2, 7] = np.nan
        ixd = BruteForceIndex.apply(metric, FuturePool.immediate_pool(), device=local)
        ixd.append_batch(ids[lo:hi], corpus2[lo:hi])
        full2 = np.empty_like(corpus)
        parts = [None] * world
        dist.all_gather_object(parts, corpus2[lo:hi])
        full2 = np.concatenate(parts)
        w2 = oracle.query_canonical(metric.ordinal, full2, ids, q2, k)
        qd2 = torch.from_numpy(q2).to(dev)
        sxd = ShardedBruteForceIndex(ixd, device=dev)
        ai, ad, ac = sxd.batch_query_device(qd2, k, st)
        torch.cuda.synchronize()
        flagged = ac.cpu().numpy() < 0
        c1 = bool(flagged[1] and flagged[2])
        good = ~flagged
        c2 = bool((ai.cpu().numpy()[good] == w2[0][good]).all())
        ok2 &= c1 and c2
        try:
            ixd.raise_pending_error()
        except Exception:
            pass
        bi, bd, bc = sxd.batch_query(qd2, k, st)
        from oracle import oracle_np as onp   # NaN distances compare by order key (every NaN is the same Float.compare value)
        c3 = bool((bi.cpu().numpy() == w2[0]).all() and (onp.float_order_key(bd.cpu().numpy()) == onp.float_order_key(w2[1])).all()
                  and (bc.cpu().numpy() == w2[2]).all())
        ok2 &= c3
        if not (c1 and c2 and c3):
            bad_rows = np.argwhere((bi.cpu().numpy() != w2[0]).any(axis=1)).ravel().tolist()
            print(f"[rank {rank}] degenerate batch: flagged_as_expected={c1} ({int(flagged.sum())} flagged: {np.argwhere(flagged).ravel().tolist()[:8]}) "
                  f"unflagged_rows_exact={c2} sync_exact={c3} bad_rows={bad_rows[:8]} counts={bc.cpu().numpy()[bad_rows[:8]].tolist()}", flush=True)
        del sxd
        ixd.close()
    # fourth: the same class on the collective route (seed bounds all-gathered over NCCL, lists all-gathered, K5 merge)
    sx = ShardedBruteForceIndex(ix, device=dev, route="allgather")
    for rep in range(2):
        si, sd, sc = sx.batch_query_device(qd, k, st)
        torch.cuda.synchronize()
        ok2 &= bool((si.cpu().numpy() == wi).all() and (sd.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
                    and (sc.cpu().numpy() == wc).all())
    ix.raise_pending_error()
    del sx

    def timed(fn, reps=20):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def route_nccl():
        dist.all_gather_into_tensor(g_ids, oi)
        dist.all_gather_into_tensor(g_dist, od)
        dist.all_gather_into_tensor(g_cnt, oc)
        merge_topk_device(g_ids, g_dist, g_cnt, k, st)

    ms_nccl = timed(route_nccl)
    ms_fused = timed(lambda: px.exchange_merge(st))
    t = torch.tensor([1 if ok else 0, 1 if ok2 else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{metric.name} n={n} d={d} b={b} k={k} world={world}: identical_to_single_shard_oracle={bool(t[0].item())} "
              f"fused_exchange_merge_identical={bool(t[1].item())} all_gather+merge={ms_nccl * 1e3:.1f}us fused={ms_fused * 1e3:.1f}us",
              flush=True)
    ok_all &= bool(t[0].item()) and bool(t[1].item())
    del px
    ix.close()
# ---- the two merge routes at the headline result shape (4096 queries x top-100), synthetic sorted lists ----
b, k = 4096, 100
px = PeerExchange(b, k, dev)
gen = torch.Generator(device=dev)
gen.manual_seed(100 + rank)
px.local.dist.copy_(torch.sort(torch.randn((b, k), generator=gen, device=dev), dim=1).values)
px.local.ids.copy_(torch.randint(0, 1 << 40, (b, k), generator=gen, device=dev))
px.local.count.fill_(k)
g_ids = torch.empty((world, b, k), dtype=torch.int64, device=dev)
g_dist = torch.empty((world, b, k), dtype=torch.float32, device=dev)
g_cnt = torch.empty((world, b), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream


def route_nccl_big():
    dist.all_gather_into_tensor(g_ids, px.local.ids)
    dist.all_gather_into_tensor(g_dist, px.local.dist)
    dist.all_gather_into_tensor(g_cnt, px.local.count)
    return merge_topk_device(g_ids, g_dist, g_cnt, k, st)


def timed_big(fn, reps=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


mi, md, mc = route_nccl_big()
fi, fd, fc = px.exchange_merge(st)
torch.cuda.synchronize()
same = bool(torch.equal(mi, fi) and torch.equal(md.view(torch.int32), fd.view(torch.int32)) and torch.equal(mc, fc))
tt = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(tt, op=dist.ReduceOp.MIN)
ms_a, ms_f = timed_big(route_nccl_big), timed_big(lambda: px.exchange_merge(st))
if rank == 0:
    print(f"b={b} k={k} world={world}: routes_identical={bool(tt.item())} all_gather+merge={ms_a * 1e3:.1f}us "
          f"fused_exchange_merge={ms_f * 1e3:.1f}us", flush=True)
ok_all &= bool(tt.item())
del px
dist.destroy_process_group()
if rank == 0:
    print("DIST_OK", ok_all, flush=True)
sys.exit(0 if ok_all else 1)
