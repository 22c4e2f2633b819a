"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bar: neighbour ids bit-identical (ties by id), distances bit-identical (the finalize kernel reproduces the oracle's
fp64-accumulate / round-once arithmetic operation for operation), so the 1e-5 relative tolerance of BASELINE.json is met
with zero slack; the tolerance is still asserted explicitly below."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # BASELINE.json north_star: distances within 1e-5 relative error in fp32


def _imports():
    from the_algorithm_b200 import _capi
    from the_algorithm_b200.ann.brute_force import BruteForceIndex, BruteForceRuntimeParams, merge_topk_device
    from the_algorithm_b200.ann.common import (ComposedQueryable, Cosine, EmbeddingProducer, EntityEmbedding, FuturePool,
                                               InnerProduct, L2, Metric, QueryableByIdImplementation,
                                               RoundRobinShardFunction, ShardedAppendable)
    return locals()


G = None


@pytest.fixture(scope="module", autouse=True)
def _g():
    global G
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    G = _imports()
    yield


def metrics():
    return [G["InnerProduct"], G["Cosine"], G["L2"]]


def make(n, d, b, seed, dup=False, perm_ids=True, scale=1.0):
    rng = np.random.default_rng(seed)
    corpus = (rng.standard_normal((n, d)) * scale / np.sqrt(d)).astype(np.float32)
    if dup and n >= 100:
        m = n // 100 + 1
        corpus[n // 2: n // 2 + m] = corpus[:m]
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    ids = (rng.permutation(n).astype(np.int64) * 13 - 17) if perm_ids else np.arange(n, dtype=np.int64)
    return corpus, ids, q


def check(metric, corpus, ids, q, k, path=0, cg=None):
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    if path:
        ix.set_option("path", path)
    if cg:
        ix.set_option("gemm_cta_group", cg)
    gi, gd, gc = ix.batch_query_with_distance(q, k)
    used = ix.stat("last_path")
    ix.close()
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, k)
    assert (gc == oc).all()
    assert (gi == oi).all(), f"ids differ at {np.argwhere(gi != oi)[:4].tolist()}"
    assert (onp.float_order_key(gd) == onp.float_order_key(od)).all()
    fin = np.isfinite(od)
    assert np.all(np.abs(gd[fin] - od[fin]) <= REL_TOL * np.abs(od[fin]))
    if path:
        assert used == path
    return used


# ------------------------------------------------------------------------------------------------ streaming scan (K2)
@pytest.mark.parametrize("mi", [0, 1, 2])
@pytest.mark.parametrize("n,d,b,k", [(1000, 16, 3, 10), (5000, 200, 9, 100), (70_001, 128, 5, 100), (3000, 100, 4, 7),
                                     (257, 36, 2, 300), (50, 8, 1, 100), (200_000, 64, 1, 200), (33, 3, 2, 5),
                                     (4097, 1, 2, 9), (9000, 999, 2, 50)])
def test_scan_path_matches_oracle(mi, n, d, b, k):
    corpus, ids, q = make(n, d, b, seed=n * 7 + d)
    check(metrics()[mi], corpus, ids, q, k, path=1)


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_scan_path_duplicates_tie_break_by_id(mi):
    corpus, ids, q = make(20_000, 200, 8, seed=5, dup=True)
    check(metrics()[mi], corpus, ids, q, 100, path=1)


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM filter (K3)
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("mi", [0, 1, 2])
@pytest.mark.parametrize("n,d,b,k", [(4096, 64, 128, 10), (50_000, 200, 300, 100), (33_333, 128, 257, 100),
                                     (20_000, 72, 64, 17), (2000, 200, 1, 100), (131_073, 40, 130, 256),
                                     (30_000, 256, 150, 100), (25_000, 285, 140, 50),
                                     (30_000, 300, 300, 100), (20_000, 384, 140, 50), (16_000, 512, 300, 20),
                                     (12_000, 637, 130, 20), (9000, 512, 7, 10)])
def test_gemm_path_matches_oracle(cg, mi, n, d, b, k):
    corpus, ids, q = make(n, d, b, seed=n + d + b, dup=(n == 33_333))
    check(metrics()[mi], corpus, ids, q, k, path=2, cg=cg)


def test_auto_path_picks_gemm_for_batches_and_scan_for_single_queries():
    corpus, ids, q = make(30_000, 200, 64, seed=1)
    assert check(G["InnerProduct"], corpus, ids, q, 100) == 2
    assert check(G["InnerProduct"], corpus, ids, q[:1], 100) == 1
    # a 300-d index streams its query tile in two K segments on the tensor-core path
    corpus, ids, q = make(20_000, 300, 40, seed=2)
    assert check(G["Cosine"], corpus, ids, q, 100) == 2
    # beyond ~640 dimensions one row tile no longer fits in shared memory: batches use the scan, correctly
    corpus, ids, q = make(6000, 700, 20, seed=3)
    assert check(G["Cosine"], corpus, ids, q, 50) == 1


# ------------------------------------------------------------------------------------------------ config 1 of BASELINE.json
def test_config1_cosine_100k_x_200_1000_queries():
    corpus, ids, q = make(100_000, 200, 1000, seed=0x5EED, perm_ids=False)
    check(G["Cosine"], corpus, ids, q, 100)


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("mi", [0, 1, 2])
def test_k_edge_cases(mi):
    metric = metrics()[mi]
    corpus, ids, q = make(300, 24, 4, seed=9)
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    i, d, c = ix.batch_query_with_distance(q, 0)                          # k == 0: success, empty
    assert i.shape == (4, 0) and c.tolist() == [0, 0, 0, 0]
    i, d, c = ix.batch_query_with_distance(q, 301)                        # k > n: n results, padded
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, 301)
    assert c.tolist() == [300] * 4 and (i == oi).all() and (i[:, 300] == -1).all() and np.isinf(d[:, 300]).all()
    with pytest.raises(G["_capi"].AnnError) as e:
        ix.batch_query_with_distance(q, -1)
    assert e.value.code == G["_capi"].ANN_ERR_NEGATIVE_K
    with pytest.raises(G["_capi"].AnnError) as e:
        ix.batch_query_with_distance(q[:, :23], 5)
    assert e.value.code == G["_capi"].ANN_ERR_DIMENSION_MISMATCH
    with pytest.raises(G["_capi"].AnnError) as e:
        ix.append_batch([1], np.zeros((1, 25), np.float32))
    assert e.value.code == G["_capi"].ANN_ERR_DIMENSION_MISMATCH
    assert ix.query(q[0], 0).result() == [] and ix.query_with_distance(q[0], -3).result() == []
    ix.close()


def test_empty_index_and_single_row():
    ix = G["BruteForceIndex"].apply(G["L2"], G["FuturePool"].immediate_pool())
    assert ix.size() == 0
    assert ix.query(np.zeros(4, np.float32), 5).result() == []
    ix.append(G["EntityEmbedding"](42, np.array([1, 2, 3, 4], np.float32))).result()
    assert ix.size() == 1
    res = ix.query_with_distance(np.array([1, 2, 3, 5], np.float32), 5).result()
    assert [(n.neighbor, n.distance.distance) for n in res] == [(42, 1.0)]
    ix.close()


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_special_values_nan_inf_zero_rows(mi):
    metric = metrics()[mi]
    rng = np.random.default_rng(3)
    corpus = rng.standard_normal((2000, 12)).astype(np.float32)
    corpus[3] = 0.0
    corpus[7, 2] = np.nan
    corpus[9, 1] = np.inf
    corpus[11, 0] = -np.inf
    corpus[13] = -0.0
    corpus[1500, 5] = np.nan
    q = rng.standard_normal((33, 12)).astype(np.float32)
    ids = rng.permutation(2000).astype(np.int64)
    used = check(metric, corpus, ids, q, 100)         # a batch goes to the tensor-core path even with special rows present:
    assert used == 2                                   # they score -1e38 there and are rescored exactly from the special list
    check(metric, corpus, ids, q, 100, path=1)        # the scan finds them through their non-finite scores
    check(metric, corpus[:1000], ids[:1000], q[:3], 1000, path=1)   # every row returned: NaN distances last, in id order
    check(metric, corpus[:1500], ids[:1500], q, 1200, path=0)       # k > 1024: exact path
    # more special rows than the list holds (256): the tensor-core path steps aside
    many = corpus.copy()
    many[100:400, 0] = np.nan
    assert check(metric, many, ids, q, 50) == 1
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, many)
    ix.set_option("path", 2)
    with pytest.raises(G["_capi"].AnnError):
        ix.batch_query_with_distance(q, 10)
    ix.close()


def test_tiny_and_huge_magnitudes():
    for scale in (1e-6, 1e6):
        corpus, ids, q = make(5000, 64, 40, seed=11, scale=scale)
        for m in metrics():
            check(m, corpus, ids, q * (1.0 if scale < 1 else 100.0), 50)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("mi", [0, 1, 2])
def test_gemm_filter_survives_worst_case_bf16_rounding(mi, cg):
    """Adversarial for the tensor-core filter: every element of the true neighbours sits just BELOW a bf16 rounding
    midpoint (the stored operand under-states each by 2^-8 relative, the worst case), every element of 300 decoys just
    ABOVE one (over-stated by as much).  Under InnerProduct the 20 victims are the exact top-20 although their approximate
    scores rank them below all 300 decoys by 0.78 % of |a||b| -- within 1.5 % of the margin the measured-residual error
    bound grants.  An error model that under-states the rounding error loses them."""
    d, e = 200, 2.0 ** -20
    victims = np.full((20, d), 1 + 2.0 ** -8 - e, np.float32)
    victims[:, -1] = 1 + 2.0 ** -7
    decoys = np.full((300, d), 1 + 2.0 ** -8 + e, np.float32)
    rng = np.random.default_rng(5)
    filler = (0.5 * rng.uniform(0.9, 1.0, (3000, d))).astype(np.float32)
    corpus = np.concatenate([filler[:1500], decoys[:150], victims, decoys[150:], filler[1500:]])
    ids = rng.permutation(corpus.shape[0]).astype(np.int64) * 3 + 1
    q = np.ones((3, d), np.float32)
    q[1] = 1 + 2.0 ** -8 - e                       # the query operand is rounded too
    q[2] = rng.uniform(0.5, 1.5, d).astype(np.float32)
    if mi == 0:   # metrics()[0] is InnerProduct, thrift ordinal 2
        oi, _, _ = oracle.query_canonical(2, corpus, ids, q[:1], 20)
        assert sorted(oi[0].tolist()) == sorted(ids[1650:1670].tolist())   # the victims really are the InnerProduct top-20
    assert check(metrics()[mi], corpus, ids, q, 20, path=2, cg=cg) == 2
    assert check(metrics()[mi], corpus, ids, q, 100, path=2, cg=cg) == 2


def test_massive_ties_are_answered_exactly_by_the_fallback():
    """100k identical rows: every row ties at rank k, far more than the bounded selector holds.  The host entry point
    re-answers such queries with the exact fallback (exact distance for every row + radix select on (distance, id)); the
    asynchronous device entry point reports them instead of guessing."""
    import torch

    corpus = np.ones((100_000, 16), np.float32)
    ids = np.random.default_rng(0).permutation(100_000).astype(np.int64)
    ix = G["BruteForceIndex"].apply(G["InnerProduct"], G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    q = np.ones((2, 16), np.float32)
    q[1, 0] = 0.5
    i, d, c = ix.batch_query_with_distance(q, 10)
    assert i.tolist() == [list(range(10))] * 2 and c.tolist() == [10, 10]
    assert d[0].tolist() == [-15.0] * 10 and d[1].tolist() == [-14.5] * 10
    assert ix.stat("exact_fallback_queries") == 2
    dev = torch.device("cuda", 0)
    oi = torch.empty((2, 10), dtype=torch.int64, device=dev)
    od = torch.empty((2, 10), dtype=torch.float32, device=dev)
    ix.query_batch_device(torch.from_numpy(q).to(dev), 10, oi, od, None)
    with pytest.raises(G["_capi"].AnnError) as e:
        ix.raise_pending_error()
    assert e.value.code == G["_capi"].ANN_ERR_CANDIDATE_OVERFLOW
    ix.set_option("device_fallback", 1)               # opt in: the device entry point synchronises and answers exactly
    ix.query_batch_device(torch.from_numpy(q).to(dev), 10, oi, od, None)
    torch.cuda.synchronize()
    ix.raise_pending_error()
    assert oi.cpu().numpy().tolist() == [list(range(10))] * 2
    ix.close()


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_degenerate_queries_nan_and_zero(mi):
    """A NaN query makes every distance NaN (all tie, order by id); a zero query does the same under Cosine."""
    metric = metrics()[mi]
    corpus, ids, q = make(7000, 20, 3, seed=77)
    q[0, 3] = np.nan
    q[1] = 0.0
    check(metric, corpus, ids, q, 25)


# ------------------------------------------------------------------------------------------------ Appendable semantics
def test_incremental_appends_grow_and_are_visible():
    metric = G["Cosine"]
    corpus, ids, q = make(30_000, 48, 20, seed=21)
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    cuts = [0, 1, 100, 1023, 1025, 7000, 30_000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ix.append_batch(ids[a:b], corpus[a:b])
        assert ix.size() == b
        gi, gd, gc = ix.batch_query_with_distance(q, 10)
        oi, od, oc = oracle.query_canonical(metric.ordinal, corpus[:b], ids[:b], q, 10)
        assert (gi == oi).all() and (gc == oc).all()
    # single-row appends through the Appendable trait are visible to the next query
    extra = q[0] * 3
    ix.append(G["EntityEmbedding"](10 ** 12, extra)).result()
    assert ix.query(extra, 1).result() == [10 ** 12]
    ix.close()


def test_generic_id_type_uses_slot_table():
    rng = np.random.default_rng(4)
    rows = rng.standard_normal((50, 6)).astype(np.float32)
    names = [f"user-{i}" for i in range(50)]
    ix = G["BruteForceIndex"].apply(G["L2"], G["FuturePool"].immediate_pool(),
                                    (G["EntityEmbedding"](n, r) for n, r in zip(names, rows)))
    got = ix.query(rows[17], 3).result()
    oi, _, _ = oracle.query_canonical(oracle.L2, rows, None, rows[17:18], 3)
    assert got == [names[j] for j in oi[0]]
    ix.close()


def test_device_pointer_entry_points_and_threaded_pool():
    import torch

    metric = G["InnerProduct"]
    corpus, ids, q = make(40_000, 200, 96, seed=31)
    dev = torch.device("cuda", 0)
    ix = G["BruteForceIndex"](metric, G["FuturePool"](4), capacity_hint=50_000)
    ix.append_batch_device(torch.from_numpy(ids).to(dev), torch.from_numpy(corpus).to(dev))
    qd = torch.from_numpy(q).to(dev)
    oi_t = torch.empty((96, 100), dtype=torch.int64, device=dev)
    od_t = torch.empty((96, 100), dtype=torch.float32, device=dev)
    oc_t = torch.empty((96,), dtype=torch.int32, device=dev)
    ix.query_batch_device(qd, 100, oi_t, od_t, oc_t, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ix.raise_pending_error()
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, 100)
    assert (oi_t.cpu().numpy() == oi).all() and (od_t.cpu().numpy().view(np.uint32) == od.view(np.uint32)).all()
    futs = [ix.query(q[j], 10) for j in range(16)]                       # FuturePool(threads): concurrent single queries
    for j, f in enumerate(futs):
        assert f.result() == oi[j, :10].tolist()
    ix.close()


# ------------------------------------------------------------------------------------------------ by-id + sharding on the GPU
def test_queryable_by_id_batches_on_device():
    rng = np.random.default_rng(8)
    corpus, ids, _ = make(20_000, 64, 1, seed=8)
    users = {u: rng.uniform(-1, 1, 64).astype(np.float32) for u in range(40)}

    class Producer(G["EmbeddingProducer"]):
        def produce_embedding(self, input):
            return users.get(input)

    ix = G["BruteForceIndex"].apply(G["Cosine"], G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    byid = G["QueryableByIdImplementation"](Producer(), ix)
    seeds = [3, 999, 7, 11]                                             # 999 has no embedding
    res = byid.batch_query_with_distance_by_id(seeds, 20, None).result()
    assert [r.seed for r in res] == [3] * 20 + [7] * 20 + [11] * 20
    for s in (3, 7, 11):
        oi, od, _ = oracle.query_canonical(oracle.COSINE, corpus, ids, users[s].reshape(1, -1), 20)
        mine = [r for r in res if r.seed == s]
        assert [r.neighbor for r in mine] == oi[0].tolist()
        assert [np.float32(r.distance.distance) for r in mine] == od[0].tolist()
    ix.close()


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_composed_queryable_over_gpu_shards_equals_single_index(mi):
    metric = metrics()[mi]
    corpus, ids, q = make(9000, 32, 37, seed=13, dup=True)
    pool = G["FuturePool"].immediate_pool()
    shards = [G["BruteForceIndex"].apply(metric, pool) for _ in range(3)]
    parts = np.array_split(np.arange(9000), 3)
    for s, p in zip(shards, parts):
        s.append_batch(ids[p], corpus[p])
    composed = G["ComposedQueryable"](shards)
    gi, gd, gc = composed.batch_query_with_distance(q, 50)
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, 50)
    assert (gi == oi).all() and (gd.view(np.uint32) == od.view(np.uint32)).all() and (gc == oc).all()
    one = composed.query_with_distance(q[0], 5).result()
    assert [n.neighbor for n in one] == oi[0, :5].tolist()
    for s in shards:
        s.close()


def test_merge_kernel_matches_oracle_merge():
    import torch

    rng = np.random.default_rng(17)
    s, b, k = 8, 19, 100
    dist = np.sort(rng.standard_normal((s, b, k)).astype(np.float32), axis=2)
    dist[:, :, ::7] = np.round(dist[:, :, ::7], 1)                      # cross-shard ties
    dist = np.sort(dist, axis=2)
    ids = rng.permutation(s * b * k).astype(np.int64).reshape(s, b, k)
    cnt = rng.integers(0, k + 1, (s, b)).astype(np.int32)
    cnt[0] = k
    dist[3, 2, :5] = np.nan
    dev = torch.device("cuda", 0)
    oi, od, oc = G["merge_topk_device"](torch.from_numpy(ids).to(dev), torch.from_numpy(dist).to(dev),
                                        torch.from_numpy(cnt).to(dev), k)
    torch.cuda.synchronize()
    oi, od, oc = oi.cpu().numpy(), od.cpu().numpy(), oc.cpu().numpy()
    for qi in range(b):
        # valid prefix of every shard list, canonical order
        ei, ed, ec = oracle.merge(ids[:, qi], dist[:, qi], cnt[:, qi], k)
        assert oc[qi] == ec and (oi[qi] == ei).all()
        assert (onp.float_order_key(od[qi]) == onp.float_order_key(ed)).all()


@pytest.mark.parametrize("world,b,k", [(1, 5, 10), (2, 37, 100), (3, 8, 1), (8, 64, 100), (16, 9, 33)])
def test_fused_exchange_merge_kernel_matches_oracle_merge(world, b, k):
    """ann_exchange_merge_device with all `world` result blocks on one device: every "rank" merges its slice of the batch
    from all local blocks and writes it to all final blocks; afterwards every final block must hold oracle.merge's answer
    (ShardApi.scala:77-85 in canonical order).  Lists are sorted by (distance, id) as the shard queries emit them, with
    cross-shard ties, duplicate (distance, id) pairs, short lists, empty lists and NaN tails."""
    import torch
    from the_algorithm_b200.ann.exchange import ResultBlock, exchange_merge_blocks, result_block_bytes, slice_of

    rng = np.random.default_rng(100 + world)
    dist = np.round(rng.standard_normal((world, b, k)).astype(np.float32), 1)     # many ties within and across shards
    ids = rng.integers(0, 50, (world, b, k)).astype(np.int64)                      # and repeated ids
    cnt = rng.integers(0, k + 1, (world, b)).astype(np.int32)
    cnt[0] = k
    if world > 1:
        cnt[1, 0] = 0
    for s in range(world):
        for qi in range(b):
            c = cnt[s, qi]
            if c > 2 and (s + qi) % 3 == 0:
                dist[s, qi, c - 2:c] = np.nan                                       # NaN sorts last (Float.compare)
            order = np.lexsort((ids[s, qi, :c], onp.float_order_key(dist[s, qi, :c])))
            dist[s, qi, :c] = dist[s, qi, :c][order]
            ids[s, qi, :c] = ids[s, qi, :c][order]
    dev = torch.device("cuda", 0)
    nb = result_block_bytes(b, k)
    assert nb == b * k * 12 + b * 4
    bufs = [torch.zeros(2 * nb + 256, dtype=torch.uint8, device=dev) for _ in range(world)]
    off = (nb + 255) // 256 * 256
    locs = [ResultBlock(bufs[s], b, k, 0) for s in range(world)]
    fins = [ResultBlock(bufs[s], b, k, off) for s in range(world)]
    for s in range(world):
        locs[s].ids.copy_(torch.from_numpy(ids[s]))
        locs[s].dist.copy_(torch.from_numpy(dist[s]))
        locs[s].count.copy_(torch.from_numpy(cnt[s]))
    for r in range(world):
        q0, q1 = slice_of(r, world, b)
        exchange_merge_blocks([x.ptr for x in locs], [x.ptr for x in fins], b, k, q0, q1 - q0, 0)
    torch.cuda.synchronize()
    for s in range(world):
        oi, od, oc = (t.cpu().numpy() for t in fins[s].tensors)
        for qi in range(b):
            ei, ed, ec = oracle.merge(ids[:, qi], dist[:, qi], cnt[:, qi], k)
            assert oc[qi] == ec
            assert (oi[qi] == ei).all(), (s, qi)
            assert (onp.float_order_key(od[qi]) == onp.float_order_key(ed)).all()


def _two_phase(shards, q_t, k, stream):
    """seed on every shard -> (same stream, so no barrier needed) -> finish on every shard against all published bounds"""
    import torch

    b = q_t.shape[0]
    dev = q_t.device
    keys = torch.empty((len(shards), b, k), dtype=torch.int32, device=dev)
    g_ids = torch.empty((len(shards), b, k), dtype=torch.int64, device=dev)
    g_dist = torch.empty((len(shards), b, k), dtype=torch.float32, device=dev)
    g_cnt = torch.empty((len(shards), b), dtype=torch.int32, device=dev)
    for s, ix in enumerate(shards):
        ix.query_seed_device(q_t, k, keys[s], stream)
    ptrs = [keys[s].data_ptr() for s in range(len(shards))]
    for s, ix in enumerate(shards):
        ix.query_finish_device(q_t, k, ptrs, g_ids[s], g_dist[s], g_cnt[s], stream)
    return keys, g_ids, g_dist, g_cnt


@pytest.mark.parametrize("mi", [0, 1, 2])
@pytest.mark.parametrize("n_shards,n_per,b", [(2, 70_001, 130), (8, 160_000, 64)])
def test_two_phase_sharded_query_shares_seed_thresholds(mi, n_shards, n_per, b):
    """ann_query_seed_device / ann_query_finish_device over the shards of one index (all on one GPU here): every shard
    publishes k bounds from its seed launch, every shard is scored against the k-th best bound of ALL shards, and the
    merged lists equal the single-index oracle answer bit for bit although the per-shard lists are cut short."""
    import torch

    metric = metrics()[mi]
    n, d, k = n_shards * n_per, 32, 100
    corpus, ids, q = make(n, d, b, seed=500 + n_shards, dup=True)
    pool = G["FuturePool"].immediate_pool()
    shards = []
    for s in range(n_shards):
        ix = G["BruteForceIndex"].apply(metric, pool)
        ix.append_batch(ids[s * n_per:(s + 1) * n_per], corpus[s * n_per:(s + 1) * n_per])
        shards.append(ix)
    dev = torch.device("cuda", 0)
    q_t = torch.from_numpy(q).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):   # the scratch and the key arrays are reused by the next batch
        keys, g_ids, g_dist, g_cnt = _two_phase(shards, q_t, k, stream)
        mi_, md_, mc_ = G["merge_topk_device"](g_ids, g_dist, g_cnt, k)
        torch.cuda.synchronize()
        for ix in shards:
            ix.raise_pending_error()
            assert ix.stat("last_path") == 2
        oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, k)
        assert (mc_.cpu().numpy() == oc).all()
        assert (mi_.cpu().numpy() == oi).all()
        assert (md_.cpu().numpy().view(np.uint32) == od.view(np.uint32)).all()
    kk = keys.cpu().numpy().view(np.uint32)
    assert (kk < 0xFF800000).all()                       # every shard seeded and published finite bounds
    cnt = g_cnt.cpu().numpy()
    assert (cnt <= k).all() and (cnt >= 0).all()
    assert (cnt.sum(axis=0) >= k).all()                  # together the lists always cover the global top-k
    # What one shard returns: real rows of ITS range with their exact distances, in canonical order, and among them every
    # row of the shard that belongs to the global answer.  (Not necessarily the shard's own exact top-k: rows beyond the
    # global bound may be missing while rows just inside its error margin are present.)
    s0 = 0
    li, ld, lc = oracle.query_canonical(metric.ordinal, corpus[:n_per], ids[:n_per], q, 4 * k)
    gi0, gd0 = g_ids[s0].cpu().numpy(), g_dist[s0].cpu().numpy()
    mine = set(ids[:n_per].tolist())
    for qi in range(b):
        c = cnt[s0, qi]
        exact = dict(zip(li[qi, :lc[qi]].tolist(), ld[qi, :lc[qi]].view(np.uint32).tolist()))
        got_ids = gi0[qi, :c].tolist()
        assert len(set(got_ids)) == c and (gi0[qi, c:] == -1).all()
        assert all(exact.get(i) == db for i, db in zip(got_ids, gd0[qi, :c].view(np.uint32).tolist()))
        keys_q = list(zip(onp.float_order_key(gd0[qi, :c]).tolist(), got_ids))
        assert keys_q == sorted(keys_q)
        assert [i for i in oi[qi].tolist() if i in mine] == [i for i in got_ids if i in set(oi[qi].tolist())]
    # sharing is worth something: one shard alone needs more chunk launches than with everyone's bounds
    if n_shards == 8:
        with_sharing = shards[0].stat("last_gemm_chunks")
        o = [torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev),
             torch.empty((b,), dtype=torch.int32, device=dev)]
        shards[0].query_batch_device(q_t, k, *o, stream)
        torch.cuda.synchronize()
        assert with_sharing == 1 and shards[0].stat("last_gemm_chunks") == 2
    for ix in shards:
        ix.close()


def test_two_phase_query_session_rules_and_unseeded_paths():
    import torch

    AnnError = G["_capi"].AnnError
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    pool = G["FuturePool"].immediate_pool()
    k = 10
    corpus, ids, q = make(3000, 24, 40, seed=77)
    ix = G["BruteForceIndex"].apply(G["L2"], pool)
    ix.append_batch(ids, corpus)
    q_t = torch.from_numpy(q).to(dev)
    outs = lambda b: (torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev),
                      torch.empty((b,), dtype=torch.int32, device=dev))
    keys = torch.zeros((40, k), dtype=torch.int32, device=dev)
    want = oracle.query_canonical(0, corpus, ids, q, k)

    def same(o, w, b=None):
        torch.cuda.synchronize()
        return all((x.cpu().numpy()[:b] == y[:b]).all() for x, y in zip(o, w))

    # a shard too small to seed (3000 rows): nothing published, finish is the ordinary tensor-core query
    o = outs(40)
    ix.query_seed_device(q_t, k, keys, stream)
    ix.query_finish_device(q_t, k, [keys.data_ptr()], *o, stream)
    assert same(o, want) and (keys.cpu().numpy() == -1).all() and ix.stat("last_path") == 2
    # a single query goes through the streaming scan: same contract
    o = outs(1)
    ix.query_seed_device(q_t[:1], k, keys, stream)
    ix.query_finish_device(q_t[:1], k, [keys.data_ptr()], *o, stream)
    assert same(o, want, 1) and ix.stat("last_path") == 1
    # no sharing requested (world = 0)
    o = outs(40)
    ix.query_seed_device(q_t, k, keys, stream)
    ix.query_finish_device(q_t, k, [], *o, stream)
    assert same(o, want)
    # finish without seed, with another shape, or after something else used the handle in between: refused
    with pytest.raises(AnnError) as e:
        ix.query_finish_device(q_t, k, [keys.data_ptr()], *o, stream)
    assert e.value.code == G["_capi"].ANN_ERR_INVALID_ARGUMENT
    ix.query_seed_device(q_t, k, keys, stream)
    with pytest.raises(AnnError):
        ix.query_finish_device(q_t[:7], k, [keys.data_ptr()], *outs(7), stream)
    ix.query_seed_device(q_t, k, keys, stream)
    ix.query_batch_device(q_t, k, *o, stream)
    with pytest.raises(AnnError):
        ix.query_finish_device(q_t, k, [keys.data_ptr()], *o, stream)
    ix.query_seed_device(q_t, k, keys, stream)
    ix.append_batch(ids[:5] + 10**6, corpus[:5])
    with pytest.raises(AnnError):
        ix.query_finish_device(q_t, k, [keys.data_ptr()], *o, stream)
    # and the handle is still good afterwards
    ix.query_seed_device(q_t, 0, keys, stream)
    ix.query_finish_device(q_t, 0, [keys.data_ptr()], *o, stream)
    ix.close()


@pytest.mark.parametrize("mi", [0, 1, 2])
@pytest.mark.parametrize("d", [1, 3, 200, 1024])
def test_metric_distance_pairs_match_oracle(mi, d):
    """Metric.distance for plain pairs (ann_distance_pairs, Metric.scala:76-158) against oracle.distance, bit for bit,
    special values included; and equal to what a query over the same rows returns."""
    metric = metrics()[mi]
    rng = np.random.default_rng(900 + d)
    n = 257
    a = (rng.standard_normal((n, d)) * 3).astype(np.float32)
    b = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    a[0], b[1] = 0.0, 0.0                       # zero vectors (Cosine: 0/0 -> NaN)
    a[2, 0], b[3, -1] = np.nan, np.inf          # non-finite entries
    a[4] = b[4]                                 # identical vectors
    a[5], b[5] = 1e-30, 1e-30                   # tiny magnitudes
    a[6], b[6] = 3e18, 3e18                     # products overflow fp32 but not fp64
    got = metric.distances(a, b)
    want = np.array([oracle.distance(metric.ordinal, a[i], b[i]) for i in range(n)], np.float32)
    assert (onp.float_order_key(got) == onp.float_order_key(want)).all()
    fin = np.isfinite(want)
    assert (got[fin].view(np.uint32) == want[fin].view(np.uint32)).all()
    one = metric.distance(a[10], b[10])
    assert isinstance(one, metric.distance_class) and np.float32(one.distance) == want[10]
    assert np.float32(metric.absolute_distance(a[10], b[10])) == want[10]
    if metric.name == "L2":
        sq = metric.distances(a[7:40], b[7:40], l2_squared=True)
        wsq = np.array([oracle.distance(0, a[i], b[i], l2_squared=1) for i in range(7, 40)], np.float32)
        assert (sq.view(np.uint32) == wsq.view(np.uint32)).all()
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    ix.append_batch(np.arange(16, dtype=np.int64), a[16:32])
    gi, gd, _ = ix.batch_query_with_distance(b[20:21], 16)
    ix.close()
    pair = metric.distances(a[16:32], np.repeat(b[20:21], 16, axis=0))
    assert (gd[0][np.argsort(gi[0])].view(np.uint32) == pair.view(np.uint32)).all()


def test_metric_util_norm_matches_oracle_and_errors():
    MetricUtil = __import__("the_algorithm_b200.ann.common", fromlist=["MetricUtil"]).MetricUtil
    AnnError = G["_capi"].AnnError
    rng = np.random.default_rng(31)
    for d in (1, 7, 200, 1024):
        rows = (rng.standard_normal((300, d)) * 20).astype(np.float32)
        rows[3] = 0.0
        rows[5, 0] = np.inf
        got = MetricUtil.norm(rows)
        want = oracle.normalize(rows)
        assert (onp.float_order_key(got) == onp.float_order_key(want)).all()
        assert (MetricUtil.norm(rows[7]).view(np.uint32) == want[7].view(np.uint32)).all()    # single vector in, single out
    assert G["L2"].distances(np.zeros((0, 5), np.float32), np.zeros((0, 5), np.float32)).shape == (0,)
    with pytest.raises(AnnError) as e:
        G["L2"].distances(np.zeros((2, 5), np.float32), np.zeros((2, 6), np.float32))
    assert e.value.code == G["_capi"].ANN_ERR_DIMENSION_MISMATCH
    with pytest.raises(AnnError):
        G["L2"].distances(np.zeros((2, 2000), np.float32), np.zeros((2, 2000), np.float32))
    with pytest.raises(AnnError):
        G["L2"].distances(np.zeros((2, 5), np.float32), np.zeros((2, 5), np.float32), device=99)


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_10m_rows():
    """At BASELINE's full size the oracle is too slow, so check size-independent properties on a 10M x 128 L2 index:
    (1) a stored row queried back finds itself first at distance exactly 0; (2) the tensor-core path and the
    streaming path (two unrelated kernels) return bit-identical lists; (3) lists are sorted under (distance, id);
    (4) appending the exact query as a new row makes it the new top-1 (visibility at scale)."""
    import torch

    dev = torch.device("cuda", 0)
    n, d, k = 10_000_000, 128, 100
    g = torch.Generator(device=dev)
    g.manual_seed(123)
    ix = G["BruteForceIndex"](G["L2"], G["FuturePool"].immediate_pool(), capacity_hint=n + 16)
    probe_rows = {}
    for c0 in range(0, n, 1_000_000):
        rows = torch.randn((1_000_000, d), generator=g, device=dev) / d ** 0.5
        ix.append_batch_device(torch.arange(c0, c0 + 1_000_000, device=dev, dtype=torch.int64), rows)
        probe_rows[c0 + 4321] = rows[4321].cpu().numpy()
    probes = np.stack(list(probe_rows.values()))
    ix.set_option("path", 1)
    i1, d1, c1 = ix.batch_query_with_distance(probes, k)
    assert i1[:, 0].tolist() == list(probe_rows.keys()) and (d1[:, 0] == 0).all()
    q = (torch.rand((160, d), generator=g, device=dev) * 2 - 1).cpu().numpy()
    si, sd, _ = ix.batch_query_with_distance(q[:8], k)
    ix.set_option("path", 2)
    gi, gd, _ = ix.batch_query_with_distance(q, k)
    assert (gi[:8] == si).all() and (gd[:8].view(np.uint32) == sd.view(np.uint32)).all()
    keys = onp.float_order_key(gd).astype(np.int64)
    assert ((keys[:, 1:] > keys[:, :-1]) | ((keys[:, 1:] == keys[:, :-1]) & (gi[:, 1:] > gi[:, :-1]))).all()
    ix.append_batch(np.array([n + 7], np.int64), q[5:6])
    gi2, gd2, _ = ix.batch_query_with_distance(q, k)
    assert gi2[5, 0] == n + 7 and gd2[5, 0] == 0 and (gi2[5, 1:] == gi[5, :-1]).all()
    ix.close()


# ------------------------------------------------------------------------------------------------ large batches, threads
def test_batch_larger_than_one_gemm_slice():
    """20k queries in one call: the tensor-core path runs them in slices of 16384 with per-query state kept for all."""
    corpus, ids, q = make(8192, 32, 20_000, seed=41)
    corpus[100:140] = corpus[0]                      # a tie group so that some queries need the margin logic
    check(G["InnerProduct"], corpus, ids, q, 10)


def test_concurrent_appends_and_queries_from_threads():
    """Appendable / Queryable are used from FuturePool threads in the reference (BruteForceIndex.scala:49,71): concurrent
    appends and queries on one handle must stay consistent; every query sees a prefix-closed set of completed appends."""
    import threading

    metric = G["L2"]
    rng = np.random.default_rng(9)
    base, ids0, q = make(20_000, 48, 32, seed=90)
    extra = (rng.standard_normal((16, 500, 48)) / 7).astype(np.float32)
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"](8))
    ix.append_batch(ids0, base)
    errors = []

    def appender(t):
        try:
            for j in range(4):
                blk = t * 4 + j
                ix.append_batch(np.arange(10 ** 9 + blk * 500, 10 ** 9 + (blk + 1) * 500, dtype=np.int64), extra[blk])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    def querier():
        try:
            for _ in range(6):
                i, d, c = ix.batch_query_with_distance(q, 20)
                assert (c == 20).all() and (np.diff(d, axis=1) >= 0).all()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=appender, args=(t,)) for t in range(4)] + [threading.Thread(target=querier) for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert ix.size() == 20_000 + 16 * 500
    # appends from different threads interleave, so compare against the oracle on the same multiset of rows
    all_ids, all_rows = ix.read_rows(0, ix.size())
    gi, gd, gc = ix.batch_query_with_distance(q, 50)
    oi, od, oc = oracle.query_canonical(metric.ordinal, all_rows, all_ids, q, 50)
    assert (gi == oi).all() and (gd.view(np.uint32) == od.view(np.uint32)).all()
    assert sorted(all_ids.tolist()) == sorted(ids0.tolist() + list(range(10 ** 9, 10 ** 9 + 8000)))
    ix.close()


# ------------------------------------------------------------------------------------------------ Updatable
@pytest.mark.parametrize("mi", [0, 1, 2])
def test_update_overwrites_in_place_and_appends_unknown_ids(mi):
    """trait Updatable (Api.scala:148-150): after update(entity) the index answers as if it had been built with the new
    embedding; the row keeps its slot and id."""
    metric = metrics()[mi]
    corpus, ids, q = make(12_000, 40, 64, seed=55)
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    rng = np.random.default_rng(56)
    pick = rng.choice(12_000, 700, replace=False)
    new_rows = (rng.standard_normal((700, 40)) / 6).astype(np.float32)
    new_rows[:3] = q[:3] * 0.9                                    # make a few updated rows clear winners
    extra_ids = np.array([10 ** 12 + 1, 10 ** 12 + 2], dtype=np.int64)
    extra_rows = (rng.standard_normal((2, 40)) / 6).astype(np.float32)
    ix.update_batch(np.concatenate([ids[pick], extra_ids]), np.concatenate([new_rows, extra_rows]))
    ix.update(G["EntityEmbedding"](int(ids[5]), corpus[5] * 2)).result()
    want_rows = corpus.copy()
    want_rows[pick] = new_rows
    want_rows[5] = corpus[5] * 2 if 5 not in pick else want_rows[5]
    if 5 in pick:
        want_rows[5] = corpus[5] * 2
    all_rows = np.concatenate([want_rows, extra_rows])
    all_ids = np.concatenate([ids, extra_ids])
    assert ix.size() == 12_002
    for path in (1, 2):
        ix.set_option("path", path)
        gi, gd, gc = ix.batch_query_with_distance(q, 30)
        oi, od, oc = oracle.query_canonical(metric.ordinal, all_rows, all_ids, q, 30)
        assert (gi == oi).all() and (gd.view(np.uint32) == od.view(np.uint32)).all()
    ix.close()


# ------------------------------------------------------------------------------------------------ query-service shell
def test_micro_batching_queryable_coalesces_concurrent_single_queries():
    """Concurrent single-vector callers behind the unchanged Queryable trait are answered in device batches."""
    from concurrent.futures import ThreadPoolExecutor

    from the_algorithm_b200.ann.service import MicroBatchingQueryable, warmup

    metric = G["Cosine"]
    corpus, ids, q = make(40_000, 64, 600, seed=71)
    ix = G["BruteForceIndex"].apply(metric, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    mb = MicroBatchingQueryable(ix, max_batch=128, max_delay_ms=2.0)
    assert warmup(mb, 64, k=100, successes=20, timeout_ms=500) >= 20
    before = mb.batches
    with ThreadPoolExecutor(32) as ex:
        futs = [ex.submit(lambda j=j: mb.query_with_distance(q[j], 10 if j % 3 else 25).result(timeout=60)) for j in range(600)]
        res = [f.result() for f in futs]
    oi10, od10, _ = oracle.query_canonical(oracle.COSINE, corpus, ids, q, 10)
    oi25, od25, _ = oracle.query_canonical(oracle.COSINE, corpus, ids, q, 25)
    for j, r in enumerate(res):
        want_i, want_d = (oi10, od10) if j % 3 else (oi25, od25)
        assert [n.neighbor for n in r] == want_i[j].tolist()
        assert [np.float32(n.distance.distance) for n in r] == want_d[j].tolist()
    assert mb.batches - before < 600 / 4                      # far fewer device calls than requests
    assert mb.query(q[0], 0).result() == []
    mb.close()
    ix.close()


def test_large_k_and_forced_exact_path():
    """min(k, size) above the bounded selectors' range (1024) is answered by the exact fallback for every query; path = 3
    forces that path, which gives an independent on-GPU cross-check of the two fast paths."""
    corpus, ids, q = make(6000, 24, 5, seed=61, dup=True)
    for m in metrics():
        ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool())
        ix.append_batch(ids, corpus)
        gi, gd, gc = ix.batch_query_with_distance(q, 3000)
        assert ix.stat("last_path") == 3
        oi, od, oc = oracle.query_canonical(m.ordinal, corpus, ids, q, 3000)
        assert (gi == oi).all() and (gd.view(np.uint32) == od.view(np.uint32)).all() and (gc == oc).all()
        ix.set_option("path", 3)
        e = ix.batch_query_with_distance(q, 40)
        ix.set_option("path", 2)
        g2 = ix.batch_query_with_distance(q, 40)
        ix.set_option("path", 1)
        s1 = ix.batch_query_with_distance(q, 40)
        assert (e[0] == g2[0]).all() and (e[0] == s1[0]).all() and (e[1].view(np.uint32) == g2[1].view(np.uint32)).all()
        ix.close()


def test_out_of_memory_is_an_error_code_not_a_crash():
    """An impossible reservation fails with ANN_ERR_OUT_OF_MEMORY and leaves no handle behind; a later index still works."""
    import ctypes

    capi = G["_capi"]
    h = ctypes.c_void_p()
    cfg = capi.AnnConfig(2, 200, 10 ** 12, 0, 0)          # 10^12 rows x 800 B: cudaMalloc refuses immediately
    rc = capi.lib().ann_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == capi.ANN_ERR_OUT_OF_MEMORY and not h.value
    assert b"failed" in capi.lib().ann_last_error()
    corpus, ids, q = make(3000, 16, 4, seed=3)
    check(G["InnerProduct"], corpus, ids, q, 10)
