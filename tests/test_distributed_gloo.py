"""world_size-2 `gloo` test of the multi-rank host class on CPU (the-algorithm_b200/ann/distributed.py): contiguous row
sharding, round-robin batch routing, replicated queries, all-gather of each rank's local top-k in the [shards][b][k]
layout, canonical merge -> identical to the single-shard answer.  The two DEVICE calls (the shard's query and the merge
kernel) are oracle-backed test doubles here (no GPU); everything between them is the product's ShardedBruteForceIndex.
On B200s the same class runs ann_query_batch_device + the fused exchange/merge kernel (tests/checks/dist_check.py, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleShard:
    """Test double for one rank's BruteForceIndex: same two methods, CPU tensors, answers from the oracle."""

    def __init__(self, metric):
        self.metric, self.ids, self.rows = metric, [], []

    def append_batch(self, ids, rows):
        self.ids.append(np.asarray(ids, np.int64))
        self.rows.append(np.asarray(rows, np.float32))

    def query_batch_device(self, queries, k, out_ids, out_dist, out_count, stream=0):
        d = queries.shape[1]
        ids = np.concatenate(self.ids) if self.ids else np.zeros((0,), np.int64)
        rows = np.concatenate(self.rows) if self.rows else np.zeros((0, d), np.float32)
        li, ld, lc = oracle.query_canonical(self.metric, rows, ids, queries.numpy(), k)
        out_ids.copy_(torch.from_numpy(li))
        out_dist.copy_(torch.from_numpy(ld))
        out_count.copy_(torch.from_numpy(lc))


def _order_keys(d):
    """Float.compare as unsigned keys (Metric.scala:17-36), vectorised: the domain the shards publish bounds in."""
    bits = np.ascontiguousarray(d, np.float32).view(np.uint32)
    keys = np.where(bits & 0x80000000, ~bits, bits | 0x80000000).astype(np.uint32)
    return np.where(np.isnan(d), np.uint32(0xFFFFFFFF), keys)


class _OracleShardTwoPhase(_OracleShard):
    """Test double for the two-phase shard query (ann_query_seed_device / ann_query_finish_device): the seed call publishes
    the k best exact distance keys of the shard's first PREFIX rows, the finish call returns the shard's top-k cut at the
    k-th smallest key of ALL shards' published bounds -- lists shorter than k, exactly what the CUDA engine hands the merge."""
    PREFIX = 200

    def _all(self, d):
        ids = np.concatenate(self.ids) if self.ids else np.zeros((0,), np.int64)
        rows = np.concatenate(self.rows) if self.rows else np.zeros((0, d), np.float32)
        return ids, rows

    def query_seed_device(self, queries, k, seed_keys, stream=0):
        ids, rows = self._all(queries.shape[1])
        _, ld, lc = oracle.query_canonical(self.metric, rows[:self.PREFIX], ids[:self.PREFIX], queries.numpy(), k)
        keys = _order_keys(ld)
        keys[np.arange(k)[None, :] >= lc[:, None]] = 0xFFFFFFFF          # fewer than k rows in the prefix: no bound
        seed_keys.copy_(torch.from_numpy(keys.view(np.int32)))
        self.pending = (int(queries.shape[0]), k)

    def query_finish_device(self, queries, k, peer_seed_key_ptrs, out_ids, out_dist, out_count, stream=0):
        import ctypes
        b = int(queries.shape[0])
        assert self.pending == (b, k)
        every = np.stack([np.frombuffer((ctypes.c_uint32 * (b * k)).from_address(int(p)), np.uint32).reshape(b, k)
                          for p in peer_seed_key_ptrs])                   # [world, b, k]
        bound = np.sort(every.transpose(1, 0, 2).reshape(b, -1), axis=1)[:, k - 1]   # k-th smallest of the union, per query
        ids, rows = self._all(queries.shape[1])
        li, ld, lc = oracle.query_canonical(self.metric, rows, ids, queries.numpy(), k)
        keep = (_order_keys(ld) <= bound[:, None]) & (np.arange(k)[None, :] < lc[:, None])
        lc = keep.sum(axis=1).astype(np.int32)                            # lists are sorted: the kept entries are a prefix
        li[~keep], ld[~keep] = -1, np.inf
        self.cut = int((lc < k).sum())
        out_ids.copy_(torch.from_numpy(li))
        out_dist.copy_(torch.from_numpy(ld))
        out_count.copy_(torch.from_numpy(lc))


def _oracle_merge(g_ids, g_dist, g_cnt, k, stream=0):
    """Test double for merge_topk_device: [S,b,k] -> [b,k] through oracle.merge."""
    b = g_ids.shape[1]
    oi, od, oc = np.empty((b, k), np.int64), np.empty((b, k), np.float32), np.empty((b,), np.int32)
    for qi in range(b):
        oi[qi], od[qi], oc[qi] = oracle.merge(g_ids[:, qi].numpy(), g_dist[:, qi].numpy(), g_cnt[:, qi].numpy(), k)
    return torch.from_numpy(oi), torch.from_numpy(od), torch.from_numpy(oc)


def _worker(rank, world, port, metric, n, d, b, k, routed, ret, two_phase=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _pkg
        _pkg.load()
        from the_algorithm_b200.ann.distributed import ShardedBruteForceIndex, shard_range

        rng = np.random.default_rng(42)                       # same stream on every rank, like bench.py
        corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
        corpus[n // 2: n // 2 + 5] = corpus[:5]              # ties across the shard boundary
        ids = rng.permutation(n).astype(np.int64)
        q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
        sx = ShardedBruteForceIndex(_OracleShardTwoPhase(metric) if two_phase else _OracleShard(metric), merge=_oracle_merge)
        assert (sx.rank, sx.world, sx.route) == (rank, world, "allgather")
        if routed:                                            # streaming appends: batches of 100 rows, round-robin
            kept = 0
            for c0 in range(0, n, 100):
                kept += int(sx.append_routed(ids[c0:c0 + 100], corpus[c0:c0 + 100]))
            assert abs(kept - ((n + 99) // 100) / world) <= 1
        else:
            lo, hi = shard_range(n, world, rank)
            sx.append_shard(ids[lo:hi], corpus[lo:hi])
        ok = True
        for _ in range(2):                                    # buffers are reused across calls
            oi, od, oc = sx.batch_query_device(torch.from_numpy(q), k)
            wi, wd, wc = oracle.query_canonical(metric, corpus, ids, q, k)
            ok &= bool((oi.numpy() == wi).all() and (od.numpy().view(np.uint32) == wd.view(np.uint32)).all()
                       and (oc.numpy() == wc).all())
        if two_phase:   # the global bound really shortened some per-shard lists (otherwise the test proves nothing new)
            c = torch.tensor([sx.local.cut])
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            ok &= int(c.item()) > 0
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("routed", [False, True])
@pytest.mark.parametrize("metric", [oracle.L2, oracle.COSINE, oracle.INNER_PRODUCT])
def test_two_rank_shard_merge_equals_single_shard(metric, routed):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, 1001, 24, 5, 16, routed, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


@pytest.mark.parametrize("metric", [oracle.L2, oracle.INNER_PRODUCT])
def test_two_rank_shared_seed_thresholds_equal_single_shard(metric):
    """The two-phase shard query: seeds published, all-gathered, every shard cut at the global bound, merged."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, 1001, 24, 5, 16, False, ret, True)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


class _FlaggingShard(_OracleShard):
    """A shard whose bounded selector gives up on query 1 (count = -1: the row is invalid) unless its exact fallback is on
    -- what the CUDA engine does for thousands of ties or a NaN query."""

    def __init__(self, metric):
        super().__init__(metric)
        self.fallback = 0

    def set_option(self, name, value):
        assert name == "device_fallback"
        self.fallback = int(value)

    def query_batch_device(self, queries, k, out_ids, out_dist, out_count, stream=0):
        super().query_batch_device(queries, k, out_ids, out_dist, out_count, stream)
        if not self.fallback:
            out_ids[1].fill_(-1)
            out_dist[1].fill_(float("inf"))
            out_count[1] = -1


def _merge_with_flags(g_ids, g_dist, g_cnt, k, stream=0):
    """merge_topk_device's contract: a query some shard flagged (count < 0) is delivered as count = -1."""
    oi, od, oc = _oracle_merge(g_ids, g_dist, torch.clamp(g_cnt, min=0), k)
    bad = (g_cnt < 0).any(dim=0)
    oi[bad], od[bad], oc[bad] = -1, float("inf"), -1
    return oi, od, oc


def _worker_slices_and_flags(rank, world, port, metric, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _pkg
        _pkg.load()
        from the_algorithm_b200.ann.distributed import ShardedBruteForceIndex, shard_range

        n, d, b, k = 700, 12, 7, 9
        rng = np.random.default_rng(3)
        corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
        ids = rng.permutation(n).astype(np.int64)
        q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
        wi, wd, wc = oracle.query_canonical(metric, corpus, ids, q, k)
        lo, hi = shard_range(n, world, rank)
        # slice delivery: this rank receives exactly its rows of the merged batch
        sx = ShardedBruteForceIndex(_OracleShard(metric), merge=_oracle_merge)
        sx.append_shard(ids[lo:hi], corpus[lo:hi])
        q0, q1 = sx.slice_range(b)
        si, sd, sc = sx.batch_query_device(torch.from_numpy(q), k, deliver="slice")
        ok = bool(si.shape[0] == q1 - q0 and (si.numpy() == wi[q0:q1]).all() and (sc.numpy() == wc[q0:q1]).all())
        # flagged rows: the asynchronous form reports them (count = -1 on every rank), the synchronous form answers exactly
        fx = ShardedBruteForceIndex(_FlaggingShard(metric), merge=_merge_with_flags)
        fx.append_shard(ids[lo:hi], corpus[lo:hi])
        ai, ad, ac = fx.batch_query_device(torch.from_numpy(q), k)
        ok &= bool(ac[1].item() == -1 and (ai[1] == -1).all() and (np.delete(ai.numpy(), 1, 0) == np.delete(wi, 1, 0)).all())
        bi, bd, bc = fx.batch_query(torch.from_numpy(q), k)
        ok &= bool((bi.numpy() == wi).all() and (bd.numpy().view(np.uint32) == wd.view(np.uint32)).all() and (bc.numpy() == wc).all())
        ok &= fx.local.fallback == 0                          # the option is switched back off afterwards
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(t.item()))
    finally:
        dist.destroy_process_group()


def test_two_rank_slice_delivery_and_flagged_rows():
    """deliver="slice" hands each rank its part of the merged batch; a query some shard cannot answer with its bounded
    selector travels through the merge as count = -1 and `batch_query` re-answers the batch with the exact fallback."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_slices_and_flags, args=(r, 2, port, oracle.COSINE, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_shard_ranges_cover_everything():
    import _pkg
    _pkg.load()
    from the_algorithm_b200.ann.distributed import route_batch, shard_range
    from the_algorithm_b200.ann.exchange import slice_of

    assert [route_batch(i, 3) for i in range(7)] == [0, 1, 2, 0, 1, 2, 0]
    for b in (1, 5, 4096):
        for world in (1, 2, 3, 8):
            sl = [slice_of(r, world, b) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == b and all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
