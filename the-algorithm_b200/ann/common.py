"""Host-side mirror of com.twitter.ann.common (the reference is Scala; there is no JVM in this image, so the
mirror above the C ABI is Python for the tests and host/scala + host/cpp for a maintainer).

Same names and argument meaning as the reference:
  Distance / L2Distance / CosineDistance / InnerProductDistance .... ann/.../common/Metric.scala:17-36
  Metric: L2, Cosine, InnerProduct, Metric.from_string ................ ann/.../common/Metric.scala:63-185
  EntityEmbedding, Queryable, Appendable, RuntimeParams,
  NeighborWithDistance, NeighborWithSeed, NeighborWithDistanceWithSeed  ann/.../common/Api.scala:9-150
  EmbeddingProducer ................................................... ann/.../common/EmbeddingProducer.scala:5-13
  QueryableById / QueryableByIdImplementation ......................... ann/.../common/QueryableById*.scala
  ShardFunction / RandomShardFunction / ShardedAppendable /
  ComposedQueryable ................................................... ann/.../common/ShardApi.scala:7-87
com.twitter.util.Future / FuturePool become concurrent.futures; Stitch becomes a plain Future.
"""
from __future__ import annotations

import math
import random
import struct
from abc import ABC, abstractmethod
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass
from typing import Any, Callable, Generic, Iterable, List, NamedTuple, Optional, Sequence, TypeVar

import numpy as np

T = TypeVar("T")
T1 = TypeVar("T1")
T2 = TypeVar("T2")


# ------------------------------------------------------------------------------------------------ futures
class FuturePool:
    """com.twitter.util.FuturePool: runs a thunk, returns a Future (BruteForceIndex.scala:49,71)."""

    def __init__(self, threads: Optional[int] = None):
        self._ex = ThreadPoolExecutor(max_workers=threads) if threads else None

    def __call__(self, fn: Callable[[], Any]) -> Future:
        if self._ex is not None:
            return self._ex.submit(fn)
        f: Future = Future()
        try:
            f.set_result(fn())
        except BaseException as e:  # failures become failed Futures, like futurePool { ... }
            f.set_exception(e)
        return f

    @staticmethod
    def immediate_pool() -> "FuturePool":
        return FuturePool(None)

    immediatePool = immediate_pool


def _float_order_key(x: float) -> int:
    """java.lang.Float.compare as an integer key (Metric.scala:22-36 use Ordering.Float.compare)."""
    f = np.float32(x)
    if math.isnan(float(f)):
        return 0xFFFFFFFF
    (b,) = struct.unpack("<I", struct.pack("<f", float(f)))
    return (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)


# ------------------------------------------------------------------------------------------------ distances
@dataclass(frozen=True)
class Distance:
    """Value class around one fp32 distance, ordered by Float.compare (Metric.scala:17-19)."""
    distance: float

    def compare(self, that: "Distance") -> int:
        a, b = _float_order_key(self.distance), _float_order_key(that.distance)
        return (a > b) - (a < b)

    def __lt__(self, that):
        return self.compare(that) < 0

    def __le__(self, that):
        return self.compare(that) <= 0

    def __gt__(self, that):
        return self.compare(that) > 0

    def __ge__(self, that):
        return self.compare(that) >= 0


class L2Distance(Distance):
    pass


class CosineDistance(Distance):
    pass


class InnerProductDistance(Distance):
    pass


class Metric:
    """sealed trait Metric[D] (Metric.scala:76-86).  `ordinal` is the thrift DistanceMetric value
    (ann_common.thrift:16-19) and is what crosses the C ABI."""

    name: str = ""
    ordinal: int = -1
    distance_class = Distance

    def from_absolute_distance(self, distance: float) -> Distance:
        return self.distance_class(float(np.float32(distance)))

    fromAbsoluteDistance = from_absolute_distance

    def distance(self, embedding1, embedding2, device: int = 0) -> Distance:
        """metric.distance(embedding1, embedding2) (Metric.scala:76-86).  Computed on the device (`ann_distance_pairs`)
        with the arithmetic of the query path, so the value is bit-identical to what queries return; there is no host
        arithmetic."""
        return self.from_absolute_distance(self.distances(np.reshape(embedding1, (1, -1)), np.reshape(embedding2, (1, -1)), device)[0])

    def distances(self, embeddings1, embeddings2, device: int = 0, l2_squared: bool = False, accum_f32: bool = False) -> np.ndarray:
        """`distance` for n pairs at once: rows of two [n, dim] arrays -> float32 [n].  `accum_f32` selects the
        sequential-fp32 accumulator convention (ANN_FLAG_ACCUM_F32)."""
        from .. import _capi

        a = np.ascontiguousarray(embeddings1, dtype=np.float32)
        b = np.ascontiguousarray(embeddings2, dtype=np.float32)
        if a.ndim != 2 or a.shape != b.shape:
            raise _capi.AnnError(_capi.ANN_ERR_DIMENSION_MISMATCH, f"embeddings differ in shape: {a.shape} vs {b.shape}")
        out = np.empty((a.shape[0],), dtype=np.float32)
        _capi.check(_capi.lib().ann_distance_pairs(self.ordinal, (_capi.ANN_FLAG_L2_SQUARED if l2_squared else 0) |
                                                   (_capi.ANN_FLAG_ACCUM_F32 if accum_f32 else 0), a.shape[1],
                                                   a.ctypes.data, b.ctypes.data, a.shape[0], out.ctypes.data, device))
        return out

    def absolute_distance(self, embedding1, embedding2, device: int = 0) -> float:
        return self.distance(embedding1, embedding2, device).distance

    absoluteDistance = absolute_distance

    def __repr__(self):
        return self.name

    @staticmethod
    def from_string(metric_name: str) -> "Metric":
        """Metric.fromString (Metric.scala:63-73); EditDistance is a string metric and is not part of this path."""
        try:
            return {"Cosine": Cosine, "L2": L2, "InnerProduct": InnerProduct}[metric_name]
        except KeyError:
            raise ValueError(f"No Metric with the name {metric_name}") from None

    fromString = from_string

    @staticmethod
    def from_thrift(ordinal: int) -> "Metric":
        return {0: L2, 1: Cosine, 2: InnerProduct}[ordinal]


class _L2(Metric):
    name, ordinal, distance_class = "L2", 0, L2Distance


class _Cosine(Metric):
    name, ordinal, distance_class = "Cosine", 1, CosineDistance


class _InnerProduct(Metric):
    name, ordinal, distance_class = "InnerProduct", 2, InnerProductDistance


class MetricUtil:
    """object MetricUtil (Metric.scala:263-290): only `norm` has callers outside Metric itself."""

    @staticmethod
    def norm(embedding, device: int = 0) -> np.ndarray:
        """MetricUtil.norm (Metric.scala:285-289): the embedding(s) scaled to unit L2 norm, on the device
        (`ann_normalize_rows`).  Accepts one vector or an [n, dim] array."""
        from .. import _capi

        e = np.ascontiguousarray(embedding, dtype=np.float32)
        rows = e.reshape(1, -1) if e.ndim == 1 else e
        out = np.empty_like(rows)
        _capi.check(_capi.lib().ann_normalize_rows(rows.shape[1], rows.ctypes.data, rows.shape[0], out.ctypes.data, device))
        return out.reshape(e.shape)


L2 = _L2()
Cosine = _Cosine()
InnerProduct = _InnerProduct()


# ------------------------------------------------------------------------------------------------ API types
class EntityEmbedding(NamedTuple):
    """case class EntityEmbedding[T](id: T, embedding: EmbeddingVector) (Api.scala:21)."""
    id: Any
    embedding: Any


class NeighborWithDistance(NamedTuple):
    neighbor: Any
    distance: Distance


class NeighborWithSeed(NamedTuple):
    seed: Any
    neighbor: Any


class NeighborWithDistanceWithSeed(NamedTuple):
    seed: Any
    neighbor: Any
    distance: Distance


class RuntimeParams:
    """trait RuntimeParams (Api.scala:90)."""


class Queryable(ABC, Generic[T]):
    """trait Queryable[T, P, D] (Api.scala:24-51)."""

    @abstractmethod
    def query(self, embedding, num_of_neighbors: int, runtime_params: RuntimeParams) -> Future:
        ...

    @abstractmethod
    def query_with_distance(self, embedding, num_of_neighbors: int, runtime_params: RuntimeParams) -> Future:
        ...

    def queryWithDistance(self, embedding, numOfNeighbors, runtimeParams):
        return self.query_with_distance(embedding, numOfNeighbors, runtimeParams)


class Appendable(ABC, Generic[T]):
    """trait Appendable[T, P, D] (Api.scala:133-145)."""

    @abstractmethod
    def append(self, entity: EntityEmbedding) -> Future:
        ...

    @abstractmethod
    def to_queryable(self) -> Queryable:
        ...

    def toQueryable(self):
        return self.to_queryable()


class EmbeddingProducer(ABC, Generic[T]):
    """trait EmbeddingProducer[T] (EmbeddingProducer.scala:5-13): id -> Option[EmbeddingVector]."""

    @abstractmethod
    def produce_embedding(self, input: T):
        ...

    def produceEmbedding(self, input):
        return self.produce_embedding(input)


def _done(value) -> Future:
    f: Future = Future()
    f.set_result(value)
    return f


# ------------------------------------------------------------------------------------------------ by id
class QueryableById(ABC):
    """trait QueryableById[T1, T2, P, D] (QueryableById.scala:17-41); Stitch[...] becomes Future[...]."""

    @abstractmethod
    def query_by_id(self, id, num_of_neighbors, runtime_params) -> Future: ...

    @abstractmethod
    def query_by_id_with_distance(self, id, num_of_neighbors, runtime_params) -> Future: ...

    @abstractmethod
    def batch_query_by_id(self, ids, num_of_neighbors, runtime_params) -> Future: ...

    @abstractmethod
    def batch_query_with_distance_by_id(self, ids, num_of_neighbors, runtime_params) -> Future: ...


class QueryableByIdImplementation(QueryableById):
    """QueryableByIdImplementation (QueryableByIdImplementation.scala:15-91): id -> embedding -> query.

    The reference's batch* methods are a Stitch.traverse of independent single-vector queries; when the wrapped
    queryable exposes `batch_query_with_distance` (the GPU index does) the whole batch goes to the device in one
    call instead.  A missing embedding or a failing lookup yields no neighbours for that id (:64, :85)."""

    def __init__(self, embedding_producer: EmbeddingProducer, queryable: Queryable):
        self.embedding_producer = embedding_producer
        self.queryable = queryable

    def query_by_id(self, id, num_of_neighbors, runtime_params) -> Future:
        emb = self.embedding_producer.produce_embedding(id)
        if emb is None:
            return _done([])
        return self.queryable.query(emb, num_of_neighbors, runtime_params)

    def query_by_id_with_distance(self, id, num_of_neighbors, runtime_params) -> Future:
        emb = self.embedding_producer.produce_embedding(id)
        if emb is None:
            return _done([])
        return self.queryable.query_with_distance(emb, num_of_neighbors, runtime_params)

    def _produce_all(self, ids):
        seeds, embs = [], []
        for i in ids:
            try:
                e = self.embedding_producer.produce_embedding(i)
            except Exception:  # .handle { case _ => List.empty }
                e = None
            if e is not None:
                seeds.append(i)
                embs.append(np.asarray(e, dtype=np.float32))
        return seeds, embs

    def batch_query_with_distance_by_id(self, ids, num_of_neighbors, runtime_params) -> Future:
        seeds, embs = self._produce_all(ids)
        out: List[NeighborWithDistanceWithSeed] = []
        if not seeds:
            return _done(out)
        batch = getattr(self.queryable, "batch_query_with_distance", None)
        if batch is not None:
            try:
                nid, dist, cnt = batch(np.stack(embs), num_of_neighbors)
                for s, seed in enumerate(seeds):
                    for j in range(int(cnt[s])):
                        out.append(NeighborWithDistanceWithSeed(seed, self.queryable.id_of(nid[s, j]),
                                                                self.queryable.metric.from_absolute_distance(dist[s, j])))
                return _done(out)
            except Exception:
                return _done([])
        for seed, e in zip(seeds, embs):
            try:
                for n in self.queryable.query_with_distance(e, num_of_neighbors, runtime_params).result():
                    out.append(NeighborWithDistanceWithSeed(seed, n.neighbor, n.distance))
            except Exception:
                pass
        return _done(out)

    def batch_query_by_id(self, ids, num_of_neighbors, runtime_params) -> Future:
        res = self.batch_query_with_distance_by_id(ids, num_of_neighbors, runtime_params).result()
        return _done([NeighborWithSeed(r.seed, r.neighbor) for r in res])

    queryById = query_by_id
    queryByIdWithDistance = query_by_id_with_distance
    batchQueryById = batch_query_by_id
    batchQueryWithDistanceById = batch_query_with_distance_by_id


# ------------------------------------------------------------------------------------------------ sharding
class ShardFunction(ABC):
    """trait ShardFunction[T] (ShardApi.scala:7-16)."""

    @abstractmethod
    def __call__(self, shards: int, entity: EntityEmbedding) -> int: ...


class RandomShardFunction(ShardFunction):
    """ShardApi.scala:21-25: Random.nextInt(shards)."""

    def __init__(self, seed: Optional[int] = None):
        self._rng = random.Random(seed)

    def __call__(self, shards: int, entity: EntityEmbedding) -> int:
        return self._rng.randrange(shards)


class RoundRobinShardFunction(ShardFunction):
    """Deterministic replacement used by the multi-GPU path: shard sizes stay within one row of each other."""

    def __init__(self):
        self._next = 0

    def __call__(self, shards: int, entity: EntityEmbedding) -> int:
        s = self._next % shards
        self._next += 1
        return s


class ShardedAppendable(Appendable):
    """ShardedAppendable (ShardApi.scala:34-48)."""

    def __init__(self, indices: Sequence[Appendable], shard_fn: ShardFunction, shards: int):
        self.indices, self.shard_fn, self.shards = list(indices), shard_fn, shards

    def append(self, entity: EntityEmbedding) -> Future:
        return self.indices[self.shard_fn(self.shards, entity)].append(entity)

    def to_queryable(self) -> Queryable:
        return ComposedQueryable([ix.to_queryable() for ix in self.indices])


class ComposedQueryable(Queryable):
    """ComposedQueryable (ShardApi.scala:58-87): query every shard, merge, take k.

    When every shard is a GPU BruteForceIndex on one device the per-shard top-k lists never leave the device:
    they are stacked [S][b][k] and merged by the K5 kernel (ann_merge_topk_device) under the (distance, id) order.
    Arbitrary Queryables are composed generically on the host with the same order (the reference's stable sort
    differs only on exact distance ties, ShardApi.scala:80-84)."""

    def __init__(self, indices: Sequence[Queryable]):
        self.indices = list(indices)

    @property
    def metric(self):
        return self.indices[0].metric

    def id_of(self, raw):
        return raw

    def _gpu_shards(self) -> bool:
        from .brute_force import BruteForceIndex

        return bool(self.indices) and all(isinstance(i, BruteForceIndex) and i.native_ids for i in self.indices) and \
            len({i.device for i in self.indices}) == 1

    def batch_query_with_distance(self, embeddings, num_of_neighbors: int):
        if not self._gpu_shards():
            raise TypeError("batch_query_with_distance needs GPU BruteForceIndex shards on one device")
        from .brute_force import merge_topk_device
        import torch

        q = np.ascontiguousarray(embeddings, dtype=np.float32)
        b, k = q.shape[0], num_of_neighbors
        dev = torch.device("cuda", self.indices[0].device)
        dq = torch.from_numpy(q).to(dev)
        s = len(self.indices)
        ids = torch.full((s, b, max(k, 1)), -1, dtype=torch.int64, device=dev)
        dist = torch.full((s, b, max(k, 1)), float("inf"), dtype=torch.float32, device=dev)
        cnt = torch.zeros((s, b), dtype=torch.int32, device=dev)
        for j, ix in enumerate(self.indices):
            if ix.size() == 0:   # an empty shard contributes an empty list (RandomShardFunction leaves shards empty early on)
                continue
            ix.query_batch_device(dq, k, ids[j], dist[j], cnt[j])
        oi, od, oc = merge_topk_device(ids, dist, cnt, k)
        torch.cuda.synchronize(dev)
        for ix in self.indices:
            ix.raise_pending_error()
        return oi.cpu().numpy()[:, :k], od.cpu().numpy()[:, :k], oc.cpu().numpy()

    def query_with_distance(self, embedding, num_of_neighbors: int, runtime_params=None) -> Future:
        if self._gpu_shards():
            ids, dist, cnt = self.batch_query_with_distance(np.asarray(embedding, dtype=np.float32).reshape(1, -1),
                                                            num_of_neighbors)
            m = self.metric
            return _done([NeighborWithDistance(int(ids[0, j]), m.from_absolute_distance(dist[0, j]))
                          for j in range(int(cnt[0]))])
        lists = [ix.query_with_distance(embedding, num_of_neighbors, runtime_params).result() for ix in self.indices]
        flat = [n for lst in lists for n in lst]
        flat.sort(key=lambda n: (_float_order_key(n.distance.distance), n.neighbor))
        return _done(flat[: max(num_of_neighbors, 0)])

    def query(self, embedding, num_of_neighbors: int, runtime_params=None) -> Future:
        return _done([n.neighbor for n in self.query_with_distance(embedding, num_of_neighbors, runtime_params).result()])
