"""Host-side mirror of com.twitter.ann.common (no GPU): names, ordering, by-id composition, sharding glue.
Fake Queryables are backed by the CPU oracle -- test doubles only."""
import math

import numpy as np
import pytest

import oracle
from the_algorithm_b200.ann.common import (ComposedQueryable, Cosine, CosineDistance, Distance, EmbeddingProducer,
                                           EntityEmbedding, FuturePool, InnerProduct, L2, L2Distance, Metric,
                                           NeighborWithDistance, Queryable, QueryableByIdImplementation,
                                           RandomShardFunction, RoundRobinShardFunction, ShardedAppendable, Appendable,
                                           _done)


def test_metric_from_string_and_ordinals():
    assert Metric.from_string("Cosine") is Cosine and Metric.from_string("L2") is L2
    assert Metric.fromString("InnerProduct") is InnerProduct
    with pytest.raises(ValueError, match="No Metric with the name"):
        Metric.from_string("Hamming")
    # thrift DistanceMetric ordinals, ann_common.thrift:16-19
    assert (L2.ordinal, Cosine.ordinal, InnerProduct.ordinal) == (0, 1, 2)
    assert Metric.from_thrift(1) is Cosine
    assert isinstance(L2.from_absolute_distance(1.5), L2Distance)


def test_distance_total_order_is_float_compare():
    vals = [float("nan"), float("inf"), 1.0, 0.0, -0.0, -1.0, float("-inf")]
    ds = sorted(CosineDistance(v) for v in vals)
    got = [d.distance for d in ds]
    assert got[0] == float("-inf") and math.isnan(got[-1])
    assert math.copysign(1, got[2]) == -1 and math.copysign(1, got[3]) == 1      # -0.0 before +0.0
    assert Distance(float("nan")).compare(Distance(float("nan"))) == 0            # every NaN equal


def test_future_pool_failure_becomes_failed_future():
    pool = FuturePool.immediate_pool()
    assert pool(lambda: 3).result() == 3
    f = pool(lambda: 1 / 0)
    assert isinstance(f.exception(), ZeroDivisionError)
    threaded = FuturePool(2)
    assert threaded(lambda: "x").result() == "x"


class OracleQueryable(Queryable, Appendable):
    """Test double: a Queryable backed by the CPU oracle."""

    def __init__(self, metric, dim):
        self.metric, self.dim = metric, dim
        self.ids, self.rows = [], []

    def append(self, entity):
        self.ids.append(entity.id)
        self.rows.append(np.asarray(entity.embedding, np.float32))
        return _done(None)

    def to_queryable(self):
        return self

    def id_of(self, raw):
        return int(raw)

    def query_with_distance(self, embedding, k, params=None):
        if not self.rows:
            return _done([])
        i, d, c = oracle.query_canonical(self.metric.ordinal, np.stack(self.rows), np.asarray(self.ids, np.int64),
                                         np.asarray(embedding, np.float32).reshape(1, -1), k)
        return _done([NeighborWithDistance(int(i[0, j]), self.metric.from_absolute_distance(d[0, j])) for j in range(c[0])])

    def query(self, embedding, k, params=None):
        return _done([n.neighbor for n in self.query_with_distance(embedding, k, params).result()])


class DictProducer(EmbeddingProducer):
    def __init__(self, table):
        self.table = table

    def produce_embedding(self, input):
        if input == "boom":
            raise RuntimeError("lookup failed")
        return self.table.get(input)


def test_queryable_by_id_semantics():
    rng = np.random.default_rng(0)
    q = OracleQueryable(Cosine, 8)
    for i in range(50):
        q.append(EntityEmbedding(i, rng.standard_normal(8)))
    table = {"a": rng.standard_normal(8).astype(np.float32), "b": rng.standard_normal(8).astype(np.float32)}
    byid = QueryableByIdImplementation(DictProducer(table), q)
    assert byid.query_by_id("missing", 5, None).result() == []                 # no embedding => empty list (:27-29)
    single = byid.query_by_id_with_distance("a", 5, None).result()
    assert [n.neighbor for n in single] == q.query(table["a"], 5).result()
    res = byid.batch_query_with_distance_by_id(["a", "missing", "boom", "b"], 3, None).result()
    assert [r.seed for r in res] == ["a"] * 3 + ["b"] * 3                      # failures and misses contribute nothing (:85)
    assert [r.neighbor for r in res[:3]] == q.query(table["a"], 3).result()
    ids_only = byid.batch_query_by_id(["b"], 3, None).result()
    assert [(r.seed, r.neighbor) for r in ids_only] == [("b", n) for n in q.query(table["b"], 3).result()]


def test_sharded_appendable_and_composed_queryable_match_single_index():
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((300, 12)).astype(np.float32)
    ids = rng.permutation(300)
    whole = OracleQueryable(L2, 12)
    shards = [OracleQueryable(L2, 12) for _ in range(4)]
    sharded = ShardedAppendable(shards, RoundRobinShardFunction(), 4)
    for i, r in zip(ids, rows):
        whole.append(EntityEmbedding(int(i), r))
        sharded.append(EntityEmbedding(int(i), r))
    assert sorted(len(s.rows) for s in shards) == [75, 75, 75, 75]
    composed = sharded.to_queryable()
    assert isinstance(composed, ComposedQueryable)
    qv = rng.standard_normal(12).astype(np.float32)
    a = whole.query_with_distance(qv, 20).result()
    b = composed.query_with_distance(qv, 20).result()
    assert [(n.neighbor, n.distance.distance) for n in a] == [(n.neighbor, n.distance.distance) for n in b]
    assert composed.query(qv, 7).result() == whole.query(qv, 7).result()
    assert composed.query(qv, 0).result() == []


def test_random_shard_function_range_and_determinism():
    f = RandomShardFunction(seed=3)
    vals = [f(5, EntityEmbedding(0, None)) for _ in range(200)]
    assert set(vals) <= set(range(5)) and len(set(vals)) == 5
    g = RandomShardFunction(seed=3)
    assert vals == [g(5, EntityEmbedding(0, None)) for _ in range(200)]
