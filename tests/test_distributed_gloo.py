"""world_size-2 `gloo` test of the multi-rank query plumbing on CPU: contiguous row sharding, replicated queries,
all-gather of each rank's local top-k in the [shards][b][k] layout, canonical merge -> identical to the single-shard
answer.  The local scan and the merge are oracle-backed test doubles here (no GPU); on a B200 the same layout feeds
ann_query_batch_device and ann_merge_topk_device (tests/test_gpu_parity.py::test_merge_kernel_*, bench.py --gpus N)."""
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


def shard_range(n, world, rank):
    """bench.py's partitioning: rank r holds rows [r*n//R, (r+1)*n//R)."""
    return rank * n // world, (rank + 1) * n // world


def _worker(rank, world, port, metric, n, d, b, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)                       # same stream on every rank, like bench.py
        corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
        corpus[n // 2: n // 2 + 5] = corpus[:5]              # ties across the shard boundary
        ids = rng.permutation(n).astype(np.int64)
        q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
        lo, hi = shard_range(n, world, rank)
        li, ld, lc = oracle.query_canonical(metric, corpus[lo:hi], ids[lo:hi], q, k)
        g_ids = torch.empty((world, b, k), dtype=torch.int64)
        g_dist = torch.empty((world, b, k), dtype=torch.float32)
        g_cnt = torch.empty((world, b), dtype=torch.int32)
        # gloo wants the list form; NCCL (bench.py) uses all_gather_into_tensor on the same [shards][b][k] buffers
        dist.all_gather(list(g_ids.unbind(0)), torch.from_numpy(li))
        dist.all_gather(list(g_dist.unbind(0)), torch.from_numpy(ld))
        dist.all_gather(list(g_cnt.unbind(0)), torch.from_numpy(lc))
        out_i = np.empty((b, k), np.int64)
        out_d = np.empty((b, k), np.float32)
        for qi in range(b):
            mi, md, mc = oracle.merge(g_ids[:, qi].numpy(), g_dist[:, qi].numpy(), g_cnt[:, qi].numpy(), k)
            out_i[qi], out_d[qi] = mi, md
        wi, wd, _ = oracle.query_canonical(metric, corpus, ids, q, k)
        ok = bool((out_i == wi).all() and (out_d.view(np.uint32) == wd.view(np.uint32)).all())
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", [oracle.L2, oracle.COSINE, oracle.INNER_PRODUCT])
def test_two_rank_shard_merge_equals_single_shard(metric):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, 1001, 24, 5, 16, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
