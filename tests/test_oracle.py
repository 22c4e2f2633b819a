"""The CPU oracle against hand-computed known answers, its numpy twin and its own two modes.

The reference ships no tests / golden vectors for ann/ (SURVEY.md F3), so the known answers below are computed by
hand from Metric.scala:88-94,119-125,150-158 under conventions C1-C7; PARITY UNPINNED is stated in DESIGN.md."""
import math

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

L2, COS, IP = oracle.L2, oracle.COSINE, oracle.INNER_PRODUCT


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------ hand-computed known answers
def test_kat_inner_product():
    # 1 - (1*4 + 2*5 + 3*6) = 1 - 32 = -31
    assert oracle.distance(IP, [1, 2, 3], [4, 5, 6]) == np.float32(-31.0)
    assert oracle.distance(IP, [0.5, -0.25], [2, 4]) == np.float32(1.0)  # dot = 0
    assert oracle.distance(IP, [0, 0], [3, 4]) == np.float32(1.0)


def test_kat_l2():
    assert oracle.distance(L2, [0, 0], [3, 4]) == np.float32(5.0)
    assert oracle.distance(L2, [0, 0], [3, 4], l2_squared=1) == np.float32(25.0)
    assert oracle.distance(L2, [1, 1, 1, 1], [1, 1, 1, 1]) == np.float32(0.0)
    assert oracle.distance(L2, [1, 2], [2, 3]) == np.float32(math.sqrt(2.0))


def test_kat_cosine():
    assert oracle.distance(COS, [1, 0], [0, 1]) == np.float32(1.0)      # orthogonal
    assert oracle.distance(COS, [1, 0], [-2, 0]) == np.float32(2.0)     # opposite
    assert oracle.distance(COS, [3, 4], [6, 8]) == np.float32(0.0)      # parallel: 50 / (5 * 10)
    # 1 - fp32(1/sqrt(2))
    assert oracle.distance(COS, [1, 0], [1, 1]) == np.float32(1.0) - np.float32(1.0 / math.sqrt(2.0))
    assert math.isnan(oracle.distance(COS, [0, 0], [1, 1]))             # zero norm => NaN (C3)


def test_kat_rounding_points():
    # C1: one rounding, fp64 accumulate.  1e8 + 1 - 1e8 is 1 in fp64, 0 in fp32-sequential.
    a, b = [1e8, 1.0, -1e8], [1.0, 1.0, 1.0]
    assert oracle.distance(IP, a, b, accum=0) == np.float32(0.0)   # 1 - 1
    assert oracle.distance(IP, a, b, accum=1) == np.float32(1.0)   # 1 - 0


def test_kat_normalize():
    """MetricUtil.norm (Metric.scala:285-289) under convention C8; hand-computed answers."""
    out = oracle.normalize([[3.0, 4.0], [0.0, -2.0], [1.0, 1.0], [0.0, 0.0]])
    assert (out[0] == np.array([0.6, 0.8], np.float32)).all()                     # fp32(3/5), fp32(4/5)
    assert (out[1] == np.array([0.0, -1.0], np.float32)).all() and not np.signbit(out[1][0])
    assert (out[2] == np.float32(1.0 / math.sqrt(2.0))).all()                      # fp32(1 / sqrt(2)) from the fp64 quotient
    assert np.isnan(out[3]).all()                                                  # zero vector: 0/0
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((64, 37)).astype(np.float32) * 50
    got = oracle.normalize(rows)
    assert (got.view(np.uint32) == onp.normalize(rows).view(np.uint32)).all()      # C and numpy restatements agree bit for bit
    assert np.allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1.0, atol=1e-6)
    # Cosine(a, b) == InnerProduct(norm(a), norm(b)) up to the roundings of the normalised copies -- the identity the
    # reference's HNSW / Faiss backends rely on (DistanceFunctionGenerator.scala:11-15)
    for i in range(0, 60, 2):
        c = oracle.distance(COS, rows[i], rows[i + 1])
        ip = oracle.distance(IP, got[i], got[i + 1])
        assert abs(float(c) - float(ip)) <= 1e-6


def test_float_compare_total_order():
    vals = np.array([np.nan, np.inf, 1.0, 0.0, -0.0, -1.0, -np.inf], dtype=np.float32)
    keys = [oracle.lib().oracle_float_order_key(float(v)) for v in vals]
    assert keys == sorted(keys, reverse=True)          # -inf < -1 < -0.0 < +0.0 < 1 < inf < NaN
    assert list(onp.float_order_key(vals)) == keys
    assert keys[3] > keys[4]                           # +0.0 above -0.0, like java.lang.Float.compare


def test_kat_topk_small():
    corpus = np.array([[1, 0], [0, 1], [1, 1], [-1, 0], [2, 0]], dtype=np.float32)
    ids = np.array([10, 11, 12, 13, 14], dtype=np.int64)
    q = np.array([[1, 0]], dtype=np.float32)
    # InnerProduct distances: 0, 1, 0, 2, -1  -> order 14(-1), 10(0), 12(0), 11(1), 13(2)
    i, d, c = oracle.query_canonical(IP, corpus, ids, q, 3)
    assert i.tolist() == [[14, 10, 12]] and d.tolist() == [[-1.0, 0.0, 0.0]] and c.tolist() == [3]
    # L2: 0, sqrt2, 1, 2, 1 -> 10, then tie (12, 14) by id, then 11
    i, d, _ = oracle.query_canonical(L2, corpus, ids, q, 4)
    assert i.tolist() == [[10, 12, 14, 11]]
    # k > n, k == 0, k < 0
    i, d, c = oracle.query_canonical(IP, corpus, ids, q, 8)
    assert c.tolist() == [5] and i[0, 5:].tolist() == [-1, -1, -1] and np.isinf(d[0, 5:]).all()
    i, d, c = oracle.query_canonical(IP, corpus, ids, q, 0)
    assert i.shape == (1, 0) and c.tolist() == [0]


# ------------------------------------------------------------------ C restatement vs numpy twin
@pytest.mark.parametrize("metric", [L2, COS, IP])
@pytest.mark.parametrize("n,d,b,k", [(500, 7, 3, 10), (3000, 200, 4, 100), (64, 128, 2, 100)])
def test_c_matches_numpy_twin(metric, n, d, b, k):
    rng = np.random.default_rng(n * 31 + d)
    corpus = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    ids = rng.permutation(n).astype(np.int64) * 3 - 7
    ci, cd, cc = oracle.query_canonical(metric, corpus, ids, q, k)
    ni, nd, nc = onp.query_canonical(metric, corpus, ids, q, k)
    assert (ci == ni).all() and (bits(cd) == bits(nd)).all() and (cc == nc).all()


@pytest.mark.parametrize("metric", [L2, COS, IP])
def test_special_values_match_twin(metric):
    rng = np.random.default_rng(3)
    corpus = rng.standard_normal((40, 6)).astype(np.float32)
    corpus[3] = 0.0
    corpus[7, 2] = np.nan
    corpus[9, 1] = np.inf
    corpus[11, 0] = -np.inf
    corpus[13] = -0.0
    q = np.array([[1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]], dtype=np.float32)
    ci, cd, _ = oracle.query_canonical(metric, corpus, None, q, 40)
    ni, nd, _ = onp.query_canonical(metric, corpus, None, q, 40)
    assert (ci == ni).all()
    assert (onp.float_order_key(cd) == onp.float_order_key(nd)).all()
    # NaN distances sort last
    assert np.isnan(cd[0, -1])


# ------------------------------------------------------------------ faithful heap emulation vs canonical
@pytest.mark.parametrize("metric", [L2, COS, IP])
def test_faithful_equals_canonical_without_ties(metric):
    rng = np.random.default_rng(11)
    corpus = (rng.standard_normal((4000, 32)) / 6).astype(np.float32)
    q = rng.uniform(-1, 1, (6, 32)).astype(np.float32)
    ids = np.arange(4000, dtype=np.int64)
    ci, cd, cc = oracle.query_canonical(metric, corpus, ids, q, 100)
    ix = oracle.FaithfulIndex(metric, 32)
    ix.append(ids, corpus)
    fi, fd, fc = ix.query(q, 100)
    assert (ci == fi).all() and (bits(cd) == bits(fd)).all() and (cc == fc).all()
    assert ix.size() == 4000


def test_faithful_differs_only_at_ties():
    """F6: the reference's tie order is heap history, not id.  The distance LISTS still agree exactly."""
    rng = np.random.default_rng(5)
    base = rng.standard_normal((50, 8)).astype(np.float32)
    corpus = np.concatenate([base, base, base])       # every distance appears three times
    ids = rng.permutation(150).astype(np.int64)
    q = rng.standard_normal((4, 8)).astype(np.float32)
    ci, cd, _ = oracle.query_canonical(IP, corpus, ids, q, 20)
    ix = oracle.FaithfulIndex(IP, 8)
    ix.append(ids, corpus)
    fi, fd, _ = ix.query(q, 20)
    assert (bits(cd) == bits(fd)).all()
    differ = ci != fi
    for qi in range(4):
        for j in np.nonzero(differ[qi])[0]:
            # a differing slot must sit inside a run of equal distances
            assert (cd[qi] == cd[qi, j]).sum() >= 2 or j == 19


def test_scala_priority_queue_trace():
    """Hand trace of mutable.PriorityQueue (Scala 2.12) with k=2 on distances [5, 5, 5] (ids a=1, b=2, c=3):
    push a -> [a]; push b -> [a, b] (b not > a, no swap); push c -> [a, b, c], size 3 > 2 -> dequeue removes a
    (root), moves c to the root, fixDown: child b not > c -> stays.  dequeueAll: c, then b -> reverse = [b, c]."""
    ix = oracle.FaithfulIndex(L2, 1)
    ix.append([1, 2, 3], np.array([[5.0], [5.0], [5.0]], dtype=np.float32))
    fi, fd, fc = ix.query(np.array([[0.0]], dtype=np.float32), 2)
    assert fi.tolist() == [[2, 3]] and fc.tolist() == [2]
    ci, _, _ = oracle.query_canonical(L2, np.array([[5.0], [5.0], [5.0]], np.float32), [1, 2, 3], [[0.0]], 2)
    assert ci.tolist() == [[1, 2]]                    # canonical: ties by id


def test_faithful_k_edge_cases():
    ix = oracle.FaithfulIndex(IP, 2)
    ix.append([7, 8], np.array([[1, 0], [0, 1]], dtype=np.float32))
    i, d, c = ix.query([[1, 0]], 0)
    assert i.shape == (1, 0) and c.tolist() == [0]
    i, d, c = ix.query([[1, 0]], 5)
    assert c.tolist() == [2] and i[0, :2].tolist() == [7, 8]
    empty = oracle.FaithfulIndex(IP, 2)
    i, d, c = empty.query([[1, 0]], 3)
    assert c.tolist() == [0]


# ------------------------------------------------------------------ shard merge (ComposedQueryable)
def test_merge_canonical_equals_single_shard():
    rng = np.random.default_rng(9)
    corpus = (rng.standard_normal((999, 16))).astype(np.float32)
    ids = rng.permutation(999).astype(np.int64)
    q = rng.standard_normal((1, 16)).astype(np.float32)
    k = 25
    whole_i, whole_d, _ = oracle.query_canonical(COS, corpus, ids, q, k)
    parts = np.array_split(np.arange(999), 4)
    si = np.stack([oracle.query_canonical(COS, corpus[p], ids[p], q, k)[0][0] for p in parts])
    sd = np.stack([oracle.query_canonical(COS, corpus[p], ids[p], q, k)[1][0] for p in parts])
    mi, md, mc = oracle.merge(si, sd, [k] * 4, k)
    assert mc == k and (mi == whole_i[0]).all() and (bits(md) == bits(whole_d[0])).all()
    ni, nd = onp.merge_canonical(si, sd, [k] * 4, k)
    assert (ni == mi).all()


def test_merge_faithful_is_stable_by_shard():
    ids = np.array([[5, 6, -1], [1, 2, -1]], dtype=np.int64)      # [S, k] slots, 2 valid per shard
    dist = np.array([[1.0, 2.0, np.inf], [1.0, 2.0, np.inf]], dtype=np.float32)
    fi, _, _ = oracle.merge(ids, dist, [2, 2], 3, faithful=True)
    assert fi.tolist() == [5, 1, 6]                  # shard order survives among ties (ShardApi.scala:80-84)
    ci, _, _ = oracle.merge(ids, dist, [2, 2], 3, faithful=False)
    assert ci.tolist() == [1, 5, 2]                  # canonical: by id


def test_accumulation_sensitivity_report():
    """(iii) of SURVEY 8(c): how often the unshipped accumulator width alone changes an id list."""
    rng = np.random.default_rng(2)
    corpus = (rng.standard_normal((20000, 200)) / np.sqrt(200)).astype(np.float32)
    q = rng.uniform(-1, 1, (8, 200)).astype(np.float32)
    a, ad, _ = oracle.query_canonical(IP, corpus, None, q, 100, accum=0)
    b, bd, _ = oracle.query_canonical(IP, corpus, None, q, 100, accum=1)
    assert np.allclose(ad, bd, rtol=1e-5, atol=1e-6)          # within the 1e-5 contract either way
    assert (np.sort(a, axis=1) == np.sort(b, axis=1)).mean() > 0.98
