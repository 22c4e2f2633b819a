"""Round-2 GPU parity tests (through the C ABI, against the CPU oracle on the same seeded inputs):
the fp32 accumulator convention (ANN_FLAG_ACCUM_F32 = oracle accum=1), the headline shape of BASELINE.json configs[1]
against the oracle on sampled queries, k beyond 4096 on the exact fallback, an empty shard under ComposedQueryable."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # BASELINE.json north_star: distances within 1e-5 relative error in fp32


def _imports():
    from the_algorithm_b200 import _capi
    from the_algorithm_b200.ann.brute_force import BruteForceIndex
    from the_algorithm_b200.ann.common import (ComposedQueryable, Cosine, EntityEmbedding, FuturePool, InnerProduct, L2,
                                               RandomShardFunction, ShardedAppendable)
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


def make(n, d, b, seed, dup=False):
    rng = np.random.default_rng(seed)
    corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    if dup and n >= 100:
        m = n // 100 + 1
        corpus[n // 2: n // 2 + m] = corpus[:m]
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    ids = rng.permutation(n).astype(np.int64) * 13 - 17
    return corpus, ids, q


def same(got, want):
    gi, gd, gc = got
    oi, od, oc = want
    assert (gc == oc).all()
    assert (gi == oi).all(), f"ids differ at {np.argwhere(gi != oi)[:4].tolist()}"
    assert (onp.float_order_key(gd) == onp.float_order_key(od)).all()


# ------------------------------------------------------------------------------------------------ accumulator convention
@pytest.mark.parametrize("mi", [0, 1, 2])
@pytest.mark.parametrize("path,n,d,b,k", [(1, 20_000, 200, 9, 100), (2, 50_000, 200, 300, 100), (2, 33_333, 128, 257, 100),
                                          (1, 3000, 37, 4, 7), (2, 20_000, 72, 64, 17), (3, 5000, 24, 3, 40)])
def test_fp32_accumulator_convention_matches_oracle_accum1(mi, path, n, d, b, k):
    """ANN_FLAG_ACCUM_F32: ids and distance bits equal the oracle's sequential-fp32 arithmetic (Metric.scala:264-269 types
    `dot` as Float) on the scan, the tensor-core filter and the exact fallback."""
    m = metrics()[mi]
    corpus, ids, q = make(n, d, b, seed=n + d + b + 1, dup=(n == 33_333))
    ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool(), accum_f32=True)
    ix.append_batch(ids, corpus)
    ix.set_option("path", path)
    got = ix.batch_query_with_distance(q, k)
    assert ix.stat("last_path") == path
    ix.close()
    same(got, oracle.query_canonical(m.ordinal, corpus, ids, q, k, accum=1))


def test_fp32_and_fp64_conventions_differ_only_where_the_oracles_differ():
    """Both conventions on the same data: each equals its own oracle, and where the two oracles agree so do the indexes."""
    corpus, ids, q = make(20_000, 200, 64, seed=99)
    m = G["InnerProduct"]
    res = {}
    for acc in (False, True):
        ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool(), accum_f32=acc)
        ix.append_batch(ids, corpus)
        res[acc] = ix.batch_query_with_distance(q, 100)
        ix.close()
    o64 = oracle.query_canonical(m.ordinal, corpus, ids, q, 100, accum=0)
    o32 = oracle.query_canonical(m.ordinal, corpus, ids, q, 100, accum=1)
    same(res[False], o64)
    same(res[True], o32)
    fin = np.isfinite(o64[1])
    assert np.all(np.abs(res[True][1][fin] - o64[1][fin]) <= REL_TOL * np.maximum(np.abs(o64[1][fin]), 1e-30))


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_metric_pairs_fp32_convention(mi):
    m = metrics()[mi]
    rng = np.random.default_rng(5)
    a = rng.standard_normal((300, 200)).astype(np.float32)
    b = rng.standard_normal((300, 200)).astype(np.float32)
    got = m.distances(a, b, accum_f32=True)
    want = np.array([oracle.distance(m.ordinal, a[i], b[i], accum=1) for i in range(300)], np.float32)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


# ------------------------------------------------------------------------------------------------ large k on the fallback
@pytest.mark.parametrize("k", [5000, 16384])
def test_exact_fallback_k_beyond_4096(k):
    """k in (4096, 16384] needs 96-192 KB of dynamic shared memory in the fallback's sort kernel (opt-in attribute)."""
    corpus, ids, q = make(20_000, 16, 2, seed=k)
    m = G["L2"]
    ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    got = ix.batch_query_with_distance(q, k)
    assert ix.stat("last_path") == 3
    ix.close()
    same(got, oracle.query_canonical(m.ordinal, corpus, ids, q, k))


# ------------------------------------------------------------------------------------------------ composed, empty shard
def test_composed_queryable_tolerates_an_empty_shard():
    """ShardedAppendable with RandomShardFunction leaves shards empty early on; the reference's ComposedQueryable just gets
    an empty list from them (ShardApi.scala:72-86)."""
    m = G["Cosine"]
    corpus, ids, q = make(3000, 32, 5, seed=8)
    shards = [G["BruteForceIndex"](m, G["FuturePool"].immediate_pool()) for _ in range(3)]
    shards[0].append_batch(ids[:1700], corpus[:1700])
    shards[2].append_batch(ids[1700:], corpus[1700:])           # shard 1 stays empty (NULL handle)
    cq = G["ComposedQueryable"](shards)
    got = cq.batch_query_with_distance(q, 20)
    same(got, oracle.query_canonical(m.ordinal, corpus, ids, q, 20))
    assert [n.neighbor for n in cq.query_with_distance(q[0], 20).result()] == got[0][0].tolist()
    for s in shards:
        s.close()


# ------------------------------------------------------------------------------------------------ headline shape vs oracle
def test_config2_shape_sampled_queries_match_oracle():
    """BASELINE.json configs[1] at full size -- InnerProduct top-100 over 10M x 200, batch 4096 on the tensor-core path --
    compared with the canonical oracle on 12 sampled queries of that batch (multi-threaded oracle, seconds)."""
    import torch

    dev = torch.device("cuda", 0)
    n, d, b, k = 10_000_000, 200, 4096, 100
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0001)
    ix = G["BruteForceIndex"](G["InnerProduct"], G["FuturePool"].immediate_pool(), capacity_hint=n)
    host = np.empty((n, d), dtype=np.float32)
    for c0 in range(0, n, 1_000_000):
        rows = torch.randn((1_000_000, d), generator=g, device=dev) / d ** 0.5
        ix.append_batch_device(torch.arange(c0, c0 + 1_000_000, device=dev, dtype=torch.int64), rows)
        host[c0:c0 + 1_000_000] = rows.cpu().numpy()
    g.manual_seed(0x5EED0002)
    q = (torch.rand((b, d), generator=g, device=dev) * 2 - 1).cpu().numpy()
    gi, gd, gc = ix.batch_query_with_distance(q, k)
    assert ix.stat("last_path") == 2
    ix.close()
    sample = np.linspace(0, b - 1, 12).astype(np.int64)
    oi, od, oc = oracle.query_canonical(2, host, None, q[sample], k)
    same((gi[sample], gd[sample], gc[sample]), (oi, od, oc))
    keys = onp.float_order_key(gd).astype(np.int64)
    assert ((keys[:, 1:] > keys[:, :-1]) | ((keys[:, 1:] == keys[:, :-1]) & (gi[:, 1:] > gi[:, :-1]))).all()


# ------------------------------------------------------------------------------------------------ one process, R shards
def _devices(shards):
    import torch

    n = torch.cuda.device_count()
    return [s % n for s in range(shards)]       # one GPU: every shard on it (same code path, events instead of NVLink)


@pytest.mark.parametrize("two_round,sliced", [(1, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("mi", [0, 1, 2])
@pytest.mark.parametrize("shards,n,d,b,k", [(2, 70_001, 200, 130, 100), (3, 50_000, 64, 40, 10), (8, 400_000, 128, 64, 100),
                                            (4, 9000, 32, 7, 100), (2, 150, 16, 5, 100), (2, 300_000, 64, 300, 50)])
def test_in_process_sharded_handle_equals_single_index(two_round, sliced, mi, shards, n, d, b, k):
    """ann_sharded_* (one process, R shards, CUDA events between the phases): bit for bit the single-index oracle answer --
    with sliced seeding (every shard seeds its slice of the batch, one bound per query), with every shard seeding every
    query (k bounds per query), and with the seed round only."""
    from the_algorithm_b200.ann.sharded import GpuShardedBruteForceIndex

    m = metrics()[mi]
    corpus, ids, q = make(n, d, b, seed=n + shards, dup=True)
    sx = GpuShardedBruteForceIndex(m, G["FuturePool"].immediate_pool(), dim=d, devices=_devices(shards))
    sx.set_option("two_round", two_round)
    sx.set_option("sliced_seeds", sliced)
    half = n // 2
    sx.append_batch(ids[:half], corpus[:half])         # two appends: every shard holds two non-adjacent row ranges
    sx.append_batch(ids[half:], corpus[half:])
    assert sx.size() == n and sum(sx.shard_sizes()) == n and max(sx.shard_sizes()) - min(sx.shard_sizes()) <= 2
    want = oracle.query_canonical(m.ordinal, corpus, ids, q, k)
    for rep in range(2):                                # repeated: buffers and events are reused
        same(sx.batch_query_with_distance(q, k), want)
    assert sx.stat("fallback_batches") == 0
    assert [x.neighbor for x in sx.query_with_distance(q[0], 5).result()] == want[0][0, :5].tolist()
    sx.close()


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_in_process_sharded_handle_answers_degenerate_batches_exactly(mi):
    """A NaN query and thousands of identical rows flag the bounded selectors of some shards: the handle must notice (count
    = -1 travels through the merge) and answer the batch again through the exact fallback -- never a silently short list."""
    from the_algorithm_b200.ann.sharded import GpuShardedBruteForceIndex

    m = metrics()[mi]
    corpus, ids, q = make(40_000, 48, 12, seed=77)
    corpus[5000:12000] = corpus[17]                     # 7000 exact duplicates: far more ties than any pool holds
    q[3] = corpus[17] * 3.0                             # ... and a query they are all nearest to
    q[7, 5] = np.nan
    sx = GpuShardedBruteForceIndex(m, G["FuturePool"].immediate_pool(), dim=48, devices=_devices(3))
    sx.append_batch(ids, corpus)
    got = sx.batch_query_with_distance(q, 100)
    assert sx.stat("fallback_batches") == 1
    same(got, oracle.query_canonical(m.ordinal, corpus, ids, q, 100))
    sx.close()


def test_in_process_sharded_handle_errors_and_empty():
    from the_algorithm_b200.ann.sharded import GpuShardedBruteForceIndex

    AnnError = G["_capi"].AnnError
    sx = GpuShardedBruteForceIndex(G["L2"], G["FuturePool"].immediate_pool(), dim=8, devices=_devices(2))
    gi, gd, gc = sx.batch_query_with_distance(np.zeros((3, 8), np.float32), 5)      # empty index: empty lists
    assert (gc == 0).all() and (gi == -1).all() and np.isinf(gd).all()
    with pytest.raises(AnnError) as e:
        sx.batch_query_with_distance(np.zeros((3, 9), np.float32), 5)
    assert e.value.code == G["_capi"].ANN_ERR_DIMENSION_MISMATCH
    with pytest.raises(AnnError):
        sx.append_batch(None, np.zeros((3, 9), np.float32))
    sx.append_batch(None, np.arange(24, dtype=np.float32).reshape(3, 8))            # ids default to the composed insertion index
    gi, gd, gc = sx.batch_query_with_distance(np.zeros((1, 8), np.float32), 5)
    assert gc[0] == 3 and gi[0, :3].tolist() == [0, 1, 2]
    assert sx.batch_query_with_distance(np.zeros((2, 8), np.float32), 0)[2].tolist() == [0, 0]
    sx.close()
    with pytest.raises(AnnError):
        GpuShardedBruteForceIndex(G["L2"], G["FuturePool"].immediate_pool(), dim=8, devices=[99])


@pytest.mark.parametrize("mi", [0, 1, 2])
def test_sliced_seed_protocol_by_hand_and_its_argument_checks(mi):
    """ann_query_seed_slice_push_device -> ann_query_filter_bounds_push_device -> ann_query_rescore_device driven shard by shard
    (two shards on one device, one stream: what ann/distributed.py and csrc/sharded.cu do across GPUs); the merged lists are
    the single-index oracle answer.  Then the session rules of the sliced calls."""
    import torch

    m = metrics()[mi]
    n, d, b, k, R = 200_000, 64, 96, 20, 2
    corpus, ids, q = make(n, d, b, seed=5, dup=True)
    dev = torch.device("cuda", 0)
    shards = []
    for r in range(R):
        lo, hi = r * n // R, (r + 1) * n // R
        ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool())
        ix.append_batch(ids[lo:hi], corpus[lo:hi])
        ix.set_option("path", 2)
        shards.append(ix)
    qd = torch.from_numpy(q).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    bounds = [torch.zeros((b,), dtype=torch.int32, device=dev) for _ in range(R)]          # every shard's [b] bound array
    kth = [torch.empty((R, b, k), dtype=torch.int32, device=dev) for _ in range(R)]        # receive buffers [source shard][b][k]
    outs = [(torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev),
             torch.empty((b,), dtype=torch.int32, device=dev)) for _ in range(R)]
    for rep in range(2):
        for r in range(R):
            q0, q1 = r * b // R, (r + 1) * b // R
            shards[r].query_seed_slice_push_device(qd, k, q0, q1 - q0, R, [t.data_ptr() for t in bounds], st)
        for r in range(R):
            shards[r].query_filter_bounds_push_device(qd, k, bounds[r].data_ptr(), R, [kth[t][r].data_ptr() for t in range(R)], st)
        for r in range(R):
            shards[r].query_rescore_device(qd, k, [kth[r][s_].data_ptr() for s_ in range(R)], *outs[r], st)
        torch.cuda.synchronize()
        assert all(int((bounds[r].view(torch.int32) == -1).sum()) == 0 for r in range(R))      # every query got a bound from its owner
        want = oracle.query_canonical(m.ordinal, corpus, ids, q, k)
        for qi in range(b):
            rows = []
            for r in range(R):
                c = int(outs[r][2][qi])
                assert 0 <= c <= k
                di = outs[r][1][qi, :c].cpu().numpy()
                ii = outs[r][0][qi, :c].cpu().numpy()
                rows += [(float(x), int(i)) for x, i in zip(di, ii)]
            rows.sort(key=lambda t: (np.inf if np.isnan(t[0]) else t[0], t[1]))        # canonical order: (distance, id)
            assert [t[1] for t in rows[:k]] == want[0][qi, :int(want[2][qi])].tolist()
    AnnError, capi = G["_capi"].AnnError, G["_capi"]
    with pytest.raises(AnnError) as e:                                                     # slice outside the batch
        shards[0].query_seed_slice_push_device(qd, k, b - 4, 8, R, [t.data_ptr() for t in bounds], st)
    assert e.value.code == capi.ANN_ERR_INVALID_ARGUMENT
    shards[0].query_seed_slice_push_device(qd, k, 0, b // 2, R, [t.data_ptr() for t in bounds], st)
    with pytest.raises(AnnError) as e:                                                     # a sliced seed needs the bounds form of the filter
        shards[0].query_filter_push_device(qd, k, [kth[0][s_].data_ptr() for s_ in range(R)], [kth[t][0].data_ptr() for t in range(R)], st)
    assert e.value.code == capi.ANN_ERR_INVALID_ARGUMENT
    shards[0].query_seed_slice_push_device(qd, k, 0, b // 2, R, [t.data_ptr() for t in bounds], st)
    with pytest.raises(AnnError) as e:                                                     # ... and cannot be finished in one round
        shards[0].query_finish_device(qd, k, [kth[0][s_].data_ptr() for s_ in range(R)], *outs[0], st)
    assert e.value.code == capi.ANN_ERR_INVALID_ARGUMENT
    torch.cuda.synchronize()
    for ix in shards:
        ix.close()


# ------------------------------------------------------------------------------------------------ the reference's on-disk format
@pytest.mark.parametrize("layout", [0, 2])
def test_thrift_directory_round_trip(tmp_path, layout):
    """ann_save_directory / ann_load_directory: BruteForceFileData as a TBinaryProtocol PersistedEmbedding stream
    (BruteForceIndex.scala:142-161, ThriftIteratorIO.scala:14-56), `_SUCCESS`, reloaded index answers identically."""
    import struct

    from the_algorithm_b200.ann.brute_force import SerializableBruteForceIndex

    m = G["Cosine"]
    corpus, ids, q = make(5000, 24, 6, seed=3)
    ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    SerializableBruteForceIndex.to_directory(ix, tmp_path / "idx", layout=layout)
    data = (tmp_path / "idx" / "BruteForceFileData").read_bytes()
    assert (tmp_path / "idx" / "_SUCCESS").exists()
    assert data[:7] == bytes([11, 0, 1, 0, 0, 0, 8]) and struct.unpack(">q", data[7:15])[0] == ids[0]     # first record's id, big-endian Long
    if layout == 0:
        assert len(data) == 5000 * (3 + 4 + 8 + 3 + 3 + 3 + 3 + 1 + 4 + 24 * 8 + 4)
    ix2 = SerializableBruteForceIndex.from_directory(tmp_path / "idx", m, G["FuturePool"].immediate_pool())
    assert ix2.size() == 5000 and ix2.dim == 24
    same(ix2.batch_query_with_distance(q, 50), ix.batch_query_with_distance(q, 50))
    same(ix2.batch_query_with_distance(q, 50), oracle.query_canonical(m.ordinal, corpus, ids, q, 50))
    ri, rr = ix2.read_rows(0, 5000)
    assert (ri == ids).all() and (rr.view(np.uint32) == corpus.view(np.uint32)).all()        # insertion order and bits preserved
    # a truncated trailing record is the end of the stream (ThriftIteratorIO.scala:42-49)
    (tmp_path / "cut").mkdir()
    (tmp_path / "cut" / "BruteForceFileData").write_bytes(data[: len(data) // 2 + 11])
    ix3 = SerializableBruteForceIndex.from_directory(tmp_path / "cut", m, G["FuturePool"].immediate_pool())
    assert 0 < ix3.size() <= 2501
    for i in (ix, ix2, ix3):
        i.close()


def test_sharded_directory_layout_and_reload_with_another_shard_count(tmp_path):
    """ShardedSerialization: `shard_<i>/BruteForceFileData` (ShardedSerialization.scala:28-38); ComposedQueryableDeserialization
    reads whatever shard directories exist (:49-66) -- here 3 written, reloaded over 2 shards, and one shard alone."""
    from the_algorithm_b200.ann.brute_force import SerializableBruteForceIndex
    from the_algorithm_b200.ann.sharded import GpuShardedBruteForceIndex

    m = G["L2"]
    corpus, ids, q = make(9000, 16, 5, seed=4)
    sx = GpuShardedBruteForceIndex(m, G["FuturePool"].immediate_pool(), dim=16, devices=_devices(3))
    sx.append_batch(ids, corpus)
    sx.to_directory(tmp_path / "sh")
    assert sorted(p.name for p in (tmp_path / "sh").iterdir()) == ["_SUCCESS", "shard_0", "shard_1", "shard_2"]
    want = oracle.query_canonical(m.ordinal, corpus, ids, q, 30)
    sx2 = GpuShardedBruteForceIndex.from_directory(tmp_path / "sh", m, G["FuturePool"].immediate_pool(), devices=_devices(2))
    assert sx2.size() == 9000 and sx2.dim == 16
    same(sx2.batch_query_with_distance(q, 30), want)
    one = SerializableBruteForceIndex.from_directory(tmp_path / "sh" / "shard_1", m, G["FuturePool"].immediate_pool())
    assert one.size() == 3000 and (one.read_rows(0, 3000)[0] == ids[3000:6000]).all()
    for i in (sx, sx2, one):
        i.close()


# ------------------------------------------------------------------------------------------------ micro-batcher, locking, storage
def test_concurrent_single_vector_calls_are_coalesced_and_exact():
    """64 native host threads issue one-vector ann_query_batch calls on one handle (what a thread-pooled query server does,
    QueryIndexThriftController.scala:39-90): every answer equals the batched answer and the oracle, and the library ran far
    fewer device batches than calls."""
    m = G["InnerProduct"]
    corpus, ids, q = make(300_000, 64, 256, seed=21)
    ix = G["BruteForceIndex"].apply(m, G["FuturePool"].immediate_pool())
    ix.append_batch(ids, corpus)
    want = oracle.query_canonical(m.ordinal, corpus, ids, q, 20)
    same(ix.batch_query_with_distance(q, 20), want)
    st = ix.loadtest(q, 20, threads=64, calls_per_thread=40, expect_ids=want[0])
    assert st["calls"] == 64 * 40 and st["mismatches"] == 0
    assert st["device_batches"] < st["calls"] / 4, st
    # switched off, every call is its own device batch -- and still exact
    ix.set_option("coalesce_max_batch", 0)
    st0 = ix.loadtest(q, 20, threads=8, calls_per_thread=10, expect_ids=want[0])
    assert st0["mismatches"] == 0 and st0["device_batches"] == 0
    ix.set_option("coalesce_max_batch", 2048)
    # Python threads through the same entry point (ctypes releases the GIL), mixed k values in flight at once
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        k = 20 if i % 3 else 7
        gi, gd, gc = ix.batch_query_with_distance(q[i % 256][None, :], k)
        return i, k, gi[0]

    w7 = oracle.query_canonical(m.ordinal, corpus, ids, q, 7)[0]
    with ThreadPoolExecutor(32) as ex:
        for i, k, gi in ex.map(one, range(600)):
            assert (gi == (want[0] if k == 20 else w7)[i % 256]).all()
    ix.close()


def test_appends_run_while_queries_run_and_storage_grows_in_place():
    """Appendable / Queryable from FuturePool threads (BruteForceIndex.scala:49,71).  The storage grows without copies
    (virtual ranges + mapped chunks): an index created with capacity_hint = 0 takes 600k rows in 150 batches while 8 threads
    query it; every answer must be the exact top-k of SOME prefix of the appended rows, and the final state equals the oracle."""
    import threading

    m = G["L2"]
    n, d, k = 600_000, 32, 10
    corpus, ids, q = make(n, d, 16, seed=31)
    ix = G["BruteForceIndex"](m, G["FuturePool"].immediate_pool(), capacity_hint=0)
    first = 4000
    ix.append_batch(ids[:first], corpus[:first])
    stop = threading.Event()
    errors, seen_sizes = [], []

    def querier(t):
        try:
            while not stop.is_set():
                gi, gd, gc = ix.batch_query_with_distance(q[t:t + 2], k)
                assert (gc == k).all()
                keys = onp.float_order_key(gd).astype(np.int64)
                assert (keys[:, 1:] >= keys[:, :-1]).all()
                seen_sizes.append(ix.size())
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=querier, args=(t,)) for t in range(8)]
    for t_ in th:
        t_.start()
    step = (n - first) // 149
    for s in range(first, n, step):
        ix.append_batch(ids[s:s + step], corpus[s:s + step])
    stop.set()
    for t_ in th:
        t_.join()
    assert not errors, errors[:1]
    assert ix.size() == n and len(set(seen_sizes)) > 3          # queries really interleaved with the appends
    assert ix.stat("mapped_bytes") < 1.6 * n * (d * 4 + 8 + 4 + 40 * 2) + (256 << 20)   # in place: no doubled footprint
    same(ix.batch_query_with_distance(q, k), oracle.query_canonical(m.ordinal, corpus, ids, q, k))
    ri, rr = ix.read_rows(0, n)
    assert (ri == ids).all() and (rr.view(np.uint32) == corpus.view(np.uint32)).all()
    ix.close()


# ------------------------------------------------------------------------------------------------ Cosine over unit rows
@pytest.mark.parametrize("path,n,d,b,k", [(1, 20_000, 200, 5, 100), (2, 50_000, 200, 200, 100), (3, 4000, 24, 3, 30)])
def test_cosine_unit_rows_is_inner_product_over_normalised_vectors(path, n, d, b, k):
    """ANN_FLAG_COSINE_UNIT_ROWS: rows stored as MetricUtil.norm(row), queries normalised, answered with InnerProduct's
    arithmetic -- how the reference's HNSW / Faiss backends do Cosine (DistanceFunctionGenerator.scala:11-30).  Oracle:
    InnerProduct over oracle-normalised corpus and queries, bit for bit; and within 1e-5 of the default Cosine arithmetic."""
    corpus, ids, q = make(n, d, b, seed=n + path)
    corpus[7] = 0.0                                                  # a zero row normalises to NaN: ordered last
    ix = G["BruteForceIndex"].apply(G["Cosine"], G["FuturePool"].immediate_pool(), cosine_unit_rows=True)
    ix.append_batch(ids, corpus)
    ix.set_option("path", path)
    got = ix.batch_query_with_distance(q, k)
    assert ix.stat("last_path") == path
    un, uq = oracle.normalize(corpus), oracle.normalize(q)
    same(got, oracle.query_canonical(oracle.INNER_PRODUCT, un, ids, uq, k))
    ri, rr = ix.read_rows(0, 64)
    assert (onp.float_order_key(rr) == onp.float_order_key(un[:64])).all()     # the stored rows ARE the unit rows
    ref = oracle.query_canonical(oracle.COSINE, corpus, ids, q, k)
    fin = np.isfinite(ref[1]) & np.isfinite(got[1])
    assert np.all(np.abs(got[1][fin] - ref[1][fin]) <= 1e-5 * np.maximum(np.abs(ref[1][fin]), 1e-3))
    ix.close()
    with pytest.raises(G["_capi"].AnnError):
        G["BruteForceIndex"].apply(G["L2"], G["FuturePool"].immediate_pool(), cosine_unit_rows=True).append_batch(ids[:4], corpus[:4])
