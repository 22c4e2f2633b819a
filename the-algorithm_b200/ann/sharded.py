"""ShardedAppendable + ComposedQueryable (ShardApi.scala:34-48, 58-87) as ONE native handle over several GPUs of one
process -- the host mirror of `ann_sharded_*` (include/b200ann.h), which is what a single-JVM caller binds
(host/scala/GpuShardedBruteForceIndex.scala).  The torchrun route (one process per GPU) lives in ann/distributed.py; both
drive the same per-device C ABI (seed -> filter -> rescore -> slice merge).

    sx = GpuShardedBruteForceIndex(Cosine, FuturePool.immediate_pool(), dim=200, devices=[0, 1, 2, 3])
    sx.append_batch(ids, rows)                         # cut into len(devices) contiguous parts, copied in parallel
    ids, dist, cnt = sx.batch_query_with_distance(q, 100)
"""
from __future__ import annotations

import ctypes
from concurrent.futures import Future
from typing import Optional, Sequence

import numpy as np

from .. import _capi
from .brute_force import BruteForceRuntimeParams, _ptr
from .common import Appendable, EntityEmbedding, FuturePool, Metric, NeighborWithDistance, Queryable


class GpuShardedBruteForceIndex(Appendable, Queryable):
    def __init__(self, metric: Metric, future_pool: FuturePool, dim: int, devices: Optional[Sequence[int]] = None,
                 n_devices: Optional[int] = None, capacity_hint: int = 0, l2_squared: bool = False, accum_f32: bool = False):
        self.metric, self.future_pool, self.dim = metric, future_pool, int(dim)
        devs = list(devices) if devices is not None else list(range(int(n_devices or 1)))
        self.devices = devs
        flags = (_capi.ANN_FLAG_L2_SQUARED if l2_squared else 0) | (_capi.ANN_FLAG_ACCUM_F32 if accum_f32 else 0)
        cfg = _capi.AnnConfig(metric.ordinal, self.dim, capacity_hint, 0, flags)
        arr = (ctypes.c_int32 * len(devs))(*devs)
        self._h = ctypes.c_void_p()
        _capi.check(_capi.lib().ann_sharded_create(ctypes.byref(cfg), arr, len(devs), ctypes.byref(self._h)))

    def close(self):
        if self._h:
            _capi.lib().ann_sharded_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ShardedAppendable -------------------------------------------------------------------------------------
    def append_batch(self, ids, rows) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise _capi.AnnError(_capi.ANN_ERR_DIMENSION_MISMATCH, f"rows must be [n, {self.dim}]")
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        if ids_a is not None and ids_a.shape[0] != rows.shape[0]:
            raise ValueError("ids and rows differ in length")
        _capi.check(_capi.lib().ann_sharded_append_batch(self._h, _ptr(ids_a), _ptr(rows), rows.shape[0]))

    def append(self, entity: EntityEmbedding) -> Future:
        return self.future_pool(lambda: self.append_batch([entity.id], np.asarray(entity.embedding, np.float32).reshape(1, -1)))

    def to_queryable(self) -> "GpuShardedBruteForceIndex":
        return self

    def size(self) -> int:
        n = ctypes.c_int64()
        _capi.check(_capi.lib().ann_sharded_size(self._h, ctypes.byref(n)))
        return int(n.value)

    def shard_sizes(self):
        out = []
        for s in range(len(self.devices)):
            h, n = ctypes.c_void_p(), ctypes.c_int64()
            _capi.check(_capi.lib().ann_sharded_shard(self._h, s, ctypes.byref(h), ctypes.byref(n)))
            out.append(int(n.value))
        return out

    # ---- ComposedQueryable -------------------------------------------------------------------------------------
    def batch_query_with_distance(self, embeddings, num_of_neighbors: int, out=None):
        """`out` = (ids [b,k] int64, dist [b,k] float32, count [b] int32) C-contiguous arrays to write into (e.g. pinned
        buffers reused from call to call); the library fills every slot, so they need no initialisation."""
        q = np.ascontiguousarray(embeddings, dtype=np.float32)
        if q.ndim != 2:
            raise ValueError("embeddings must be [b, dim]")
        b, k = q.shape[0], int(num_of_neighbors)
        if k < 0:
            raise _capi.AnnError(_capi.ANN_ERR_NEGATIVE_K, "numOfNeighbours < 0")
        if out is not None:
            out_ids, out_dist, out_cnt = out
            if (out_ids.shape != (b, k) or out_dist.shape != (b, k) or out_cnt.shape != (b,) or out_ids.dtype != np.int64
                    or out_dist.dtype != np.float32 or out_cnt.dtype != np.int32
                    or not (out_ids.flags.c_contiguous and out_dist.flags.c_contiguous and out_cnt.flags.c_contiguous)):
                raise ValueError("out must be C-contiguous (int64 [b,k], float32 [b,k], int32 [b])")
        else:
            out_ids = np.full((b, k), -1, dtype=np.int64)
            out_dist = np.full((b, k), np.inf, dtype=np.float32)
            out_cnt = np.zeros(b, dtype=np.int32)
        _capi.check(_capi.lib().ann_sharded_query_batch(self._h, _ptr(q), b, q.shape[1], k, _ptr(out_ids), _ptr(out_dist), _ptr(out_cnt)))
        return out_ids, out_dist, out_cnt

    def id_of(self, raw):
        return int(raw)

    def query_with_distance(self, embedding, num_of_neighbors: int, runtime_params=BruteForceRuntimeParams) -> Future:
        def run():
            if num_of_neighbors <= 0:
                return []
            ids, dist, cnt = self.batch_query_with_distance(np.asarray(embedding, np.float32).reshape(1, -1), num_of_neighbors)
            return [NeighborWithDistance(int(ids[0, j]), self.metric.from_absolute_distance(dist[0, j])) for j in range(int(cnt[0]))]
        return self.future_pool(run)

    def query(self, embedding, num_of_neighbors: int, runtime_params=BruteForceRuntimeParams) -> Future:
        def run():
            return [n.neighbor for n in self.query_with_distance(embedding, num_of_neighbors, runtime_params).result()]
        return self.future_pool(run)

    # ---- ShardedSerialization / ComposedQueryableDeserialization (ShardedSerialization.scala:17-66) ---------------
    def to_directory(self, directory, id_format: int = _capi.ANN_ID_INT64_BE, layout: int = _capi.ANN_LAYOUT_FLOAT_TENSOR) -> None:
        """`shard_<i>/BruteForceFileData` per shard, the reference's thrift stream (csrc/persist.cu), plus `_SUCCESS`."""
        import os

        _capi.check(_capi.lib().ann_sharded_save_directory(self._h, os.fsencode(str(directory)), id_format, layout))

    toDirectory = to_directory

    @classmethod
    def from_directory(cls, directory, metric: Metric, future_pool: FuturePool, devices: Sequence[int], dim: int = 0,
                       id_format: int = _capi.ANN_ID_AUTO) -> "GpuShardedBruteForceIndex":
        """Loads `shard_<i>/` directories written with ANY number of shards (or one unsharded index directory) into a
        composed handle over `devices`; rows are re-dealt over the devices."""
        import os

        self = cls.__new__(cls)
        self.metric, self.future_pool, self.devices = metric, future_pool, list(devices)
        cfg = _capi.AnnConfig(metric.ordinal, int(dim), 0, 0, 0)
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self._h = ctypes.c_void_p()
        _capi.check(_capi.lib().ann_sharded_load_directory(ctypes.byref(cfg), os.fsencode(str(directory)), id_format, arr,
                                                           len(self.devices), ctypes.byref(self._h)))
        self.dim = self.stat("dim")     # the composed handle's dimension (not summed over the shards)
        return self

    fromDirectory = from_directory

    def set_option(self, name: str, value: int) -> None:
        _capi.check(_capi.lib().ann_sharded_set_option(self._h, name.encode(), int(value)))

    def stat(self, name: str) -> int:
        v = ctypes.c_int64()
        _capi.check(_capi.lib().ann_sharded_get_stat(self._h, name.encode(), ctypes.byref(v)))
        return int(v.value)
