"""Host-side mirror of com.twitter.ann.brute_force.BruteForceIndex over the CUDA engine
(ann/src/main/scala/com/twitter/ann/brute_force/BruteForceIndex.scala:24-92).

    BruteForceIndex.apply(metric, future_pool, initial_embeddings)   <- object BruteForceIndex.apply   (:29-37)
    index.append(EntityEmbedding(id, embedding)) -> Future[None]     <- append                          (:48-52)
    index.to_queryable()                                             <- toQueryable                     (:54)
    index.query(embedding, k, BruteForceRuntimeParams)               <- query                           (:56-64)
    index.query_with_distance(embedding, k, BruteForceRuntimeParams) <- queryWithDistance               (:66-91)
plus what a device-resident index adds: append_batch / batch_query_with_distance (whole batches in one call) and
the raw device-pointer entry points used by the multi-GPU merge and by bench.py.

Every compute call goes through the C ABI (include/b200ann.h); nothing here does distance arithmetic or
selection on the host, and the module raises if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes
import threading
from concurrent.futures import Future
from typing import Iterable, Iterator, List, Optional

import numpy as np

from .. import _capi
from .common import (Appendable, EntityEmbedding, FuturePool, Metric, NeighborWithDistance, Queryable, RuntimeParams)


class _BruteForceRuntimeParams(RuntimeParams):
    """object BruteForceRuntimeParams extends RuntimeParams (BruteForceIndex.scala:24): no knobs, search is exact."""

    def __repr__(self):
        return "BruteForceRuntimeParams"


BruteForceRuntimeParams = _BruteForceRuntimeParams()

_I64_MIN, _I64_MAX = -(2 ** 63), 2 ** 63 - 1


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class BruteForceIndex(Appendable, Queryable):
    DataFileName = "BruteForceFileData"
    _PENDING_FLUSH = 4096

    def __init__(self, metric: Metric, future_pool: FuturePool, device: int = 0, capacity_hint: int = 0,
                 l2_squared: bool = False, shadow: bool = True, accum_f32: bool = False, cosine_unit_rows: bool = False):
        self.metric = metric
        self.future_pool = future_pool
        self.device = device
        self.dim: Optional[int] = None
        self._h = ctypes.c_void_p()
        # accum_f32: distances follow the sequential-fp32 accumulator convention (ANN_FLAG_ACCUM_F32, include/b200ann.h)
        self._cfg = dict(capacity_hint=capacity_hint, flags=(_capi.ANN_FLAG_L2_SQUARED if l2_squared else 0) |
                         (0 if shadow else _capi.ANN_FLAG_NO_SHADOW) | (_capi.ANN_FLAG_ACCUM_F32 if accum_f32 else 0) |
                         (_capi.ANN_FLAG_COSINE_UNIT_ROWS if cosine_unit_rows else 0))
        self._slots = None
        self._slots_version = -1
        self._version = 0          # bumped by every successful append: the id -> slot map is rebuilt when it lags
        # `_lock` guards the host-side buffer of single-row appends; `_append_lock` serialises appends / updates (id
        # bookkeeping + the native call).  Queries hold NEITHER during the native call: the library lets appends proceed while
        # queries run and combines concurrent small queries itself (include/b200ann.h).
        self._lock = threading.RLock()
        self._append_lock = threading.RLock()
        self._pending_ids: List = []
        self._pending_rows: List[np.ndarray] = []
        # generic id type T: ids that are not int64 live in a host table, the device sees the insertion slot
        self.native_ids = True
        self._id_table: Optional[List] = None
        self._n = 0

    # ---- construction -------------------------------------------------------------------------------------
    @staticmethod
    def apply(metric: Metric, future_pool: FuturePool, initial_embeddings: Iterable[EntityEmbedding] = (), *,
              device: int = 0, capacity_hint: int = 0, l2_squared: bool = False, shadow: bool = True,
              accum_f32: bool = False, cosine_unit_rows: bool = False) -> "BruteForceIndex":
        ix = BruteForceIndex(metric, future_pool, device, capacity_hint, l2_squared, shadow, accum_f32, cosine_unit_rows)
        ids, rows = [], []
        for e in initial_embeddings:
            ids.append(e.id)
            rows.append(np.asarray(e.embedding, dtype=np.float32))
        if rows:
            ix.append_batch(ids, np.stack(rows))
        return ix

    def _ensure(self, dim: int):
        if self._h:
            if dim != self.dim:
                raise _capi.AnnError(_capi.ANN_ERR_DIMENSION_MISMATCH,
                                     f"embedding dimension {dim} != index dimension {self.dim}")
            return
        cfg = _capi.AnnConfig(self.metric.ordinal, dim, self._cfg["capacity_hint"], self.device, self._cfg["flags"])
        _capi.check(_capi.lib().ann_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self.dim = dim

    def close(self):
        with self._lock, self._append_lock:
            if self._h:
                _capi.lib().ann_destroy(self._h)
                self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Appendable ----------------------------------------------------------------------------------------
    def append(self, entity: EntityEmbedding) -> Future:
        """One row.  Rows are gathered on the host and reach the device as one batch at the next query / size /
        flush (or every 4096 rows), so a query still observes every append that returned before it was issued."""
        def run():
            with self._lock:
                self._pending_ids.append(entity.id)
                self._pending_rows.append(np.asarray(entity.embedding, dtype=np.float32))
                if len(self._pending_rows) >= self._PENDING_FLUSH:
                    self.flush()
        return self.future_pool(run)

    def flush(self):
        with self._lock:
            if not self._pending_rows:
                return
            ids, rows = self._pending_ids, self._pending_rows
            dims = {r.shape[-1] for r in rows}
            if len(dims) != 1:   # nothing is dropped: the rows stay buffered, every later flush reports the same error
                raise _capi.AnnError(_capi.ANN_ERR_DIMENSION_MISMATCH, "appended embeddings differ in dimension")
            self.append_batch(ids, np.stack(rows))      # raises before the buffers are released
            self._pending_ids, self._pending_rows = [], []

    def _device_ids(self, ids, n: int):
        """Device ids of a batch + the entries its success adds to the host id table (committed by the caller only after
        the C call succeeded, so a failed append cannot desynchronise slot and id)."""
        if ids is None:
            dev = np.arange(self._n, self._n + n, dtype=np.int64)
            return dev, (dev.tolist() if self._id_table is not None else None)
        if isinstance(ids, np.ndarray) and ids.dtype.kind in "iu" and self._id_table is None:
            return np.ascontiguousarray(ids, dtype=np.int64), None
        ids = list(ids)
        if self._id_table is None and all(isinstance(i, (int, np.integer)) and _I64_MIN <= int(i) <= _I64_MAX for i in ids):
            return np.asarray(ids, dtype=np.int64), None
        # generic T: slot table.  Ties then break by insertion slot (documented in DESIGN.md).
        if self._id_table is None and self._n:
            raise TypeError("cannot mix native int64 ids with generic ids in one index")
        return np.arange(self._n, self._n + n, dtype=np.int64), ids

    def append_batch(self, ids, rows) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2:
            raise ValueError("rows must be [n, dim]")
        n = rows.shape[0]
        with self._append_lock:
            if n == 0:
                return
            self._ensure(rows.shape[1])
            dev_ids, table_add = self._device_ids(ids, n)
            if dev_ids.shape[0] != n or (table_add is not None and len(table_add) != n):
                raise ValueError("ids and rows differ in length")
            _capi.check(_capi.lib().ann_append_batch(self._h, _ptr(dev_ids), _ptr(rows), n))
            if table_add is not None:
                if self._id_table is None:
                    self._id_table = []
                    self.native_ids = False
                self._id_table.extend(table_add)
            self._n += n
            self._version += 1

    def append_batch_device(self, ids_t, rows_t, stream: int = 0) -> None:
        """rows_t: CUDA float32 [n, dim] tensor, ids_t: CUDA int64 [n] tensor (or None) on this index's device."""
        with self._append_lock:
            n = int(rows_t.shape[0])
            if n == 0:
                return
            self._ensure(int(rows_t.shape[1]))
            if self._id_table is not None:
                raise TypeError("device appends need native int64 ids")
            _capi.check(_capi.lib().ann_append_batch_device(
                self._h, None if ids_t is None else ctypes.c_void_p(ids_t.data_ptr()), ctypes.c_void_p(rows_t.data_ptr()),
                n, ctypes.c_void_p(stream)))
            self._n += n
            self._version += 1

    # ---- Updatable (Api.scala:148-150) ---------------------------------------------------------------------------
    def _slot_map(self):
        """id -> insertion slot of the FIRST row carrying that id, rebuilt whenever rows were appended since it was built."""
        if self._slots is None or self._slots_version != self._version:
            if self._id_table is not None:
                it = self._id_table
            else:
                ids, _ = self.read_rows(0, self._n) if self._n else (np.zeros(0, np.int64), None)
                it = (int(i) for i in ids)
            slots = {}
            for s, i in enumerate(it):
                slots.setdefault(i, s)
            self._slots, self._slots_version = slots, self._version
        return self._slots

    def update_batch(self, ids, rows) -> None:
        """Overwrite the embeddings of existing ids in place; unknown ids are appended (Hnsw.update semantics,
        hnsw/Hnsw.scala:149-182).  With duplicate ids in the index the first inserted one is updated."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        with self._append_lock:
            self.flush()
            smap = self._slot_map() if self._n else {}
            ids = list(ids)
            key = (lambda i: i) if self._id_table is not None else int
            known = [j for j, i in enumerate(ids) if key(i) in smap]
            known_set = set(known)
            fresh = [j for j in range(len(ids)) if j not in known_set]
            if known:
                slots = np.asarray([smap[key(ids[j])] for j in known], dtype=np.int64)
                sub = np.ascontiguousarray(rows[known])
                if sub.shape[1] != self.dim:
                    raise _capi.AnnError(_capi.ANN_ERR_DIMENSION_MISMATCH, "embedding dimension != index dimension")
                _capi.check(_capi.lib().ann_update_batch(self._h, _ptr(slots), _ptr(sub), len(known)))
            if fresh:
                self.append_batch([ids[j] for j in fresh], rows[fresh])

    def update(self, entity: EntityEmbedding) -> Future:
        return self.future_pool(lambda: self.update_batch([entity.id], np.asarray(entity.embedding, np.float32).reshape(1, -1)))

    def to_queryable(self) -> "BruteForceIndex":
        return self

    def size(self) -> int:
        self.flush()
        if not self._h:
            return 0
        n = ctypes.c_int64()
        _capi.check(_capi.lib().ann_size(self._h, ctypes.byref(n)))
        return int(n.value)

    # ---- Queryable -----------------------------------------------------------------------------------------
    def id_of(self, raw):
        raw = int(raw)
        return raw if self._id_table is None else self._id_table[raw]

    def batch_query_with_distance(self, embeddings, num_of_neighbors: int, out=None):
        """b queries in one call.  Returns (ids [b,k] int64 (device ids), distances [b,k] fp32, counts [b]).
        `out` = (ids, dist, counts) lets the caller supply the result buffers (e.g. pinned host memory), which the C ABI
        fills completely; otherwise fresh arrays are allocated."""
        q = np.ascontiguousarray(embeddings, dtype=np.float32)
        if q.ndim != 2:
            raise ValueError("embeddings must be [b, dim]")
        b, k = q.shape[0], int(num_of_neighbors)
        self.flush()
        kk = max(k, 0)
        if out is not None and self._h and k > 0:
            out_ids, out_dist, out_cnt = out
            assert out_ids.shape == (b, kk) and out_ids.dtype == np.int64 and out_ids.flags.c_contiguous
            assert out_dist.shape == (b, kk) and out_dist.dtype == np.float32 and out_dist.flags.c_contiguous
            assert out_cnt.shape == (b,) and out_cnt.dtype == np.int32
        else:
            out_ids = np.full((b, kk), -1, dtype=np.int64)
            out_dist = np.full((b, kk), np.inf, dtype=np.float32)
            out_cnt = np.zeros(b, dtype=np.int32)
        if k < 0:
            raise _capi.AnnError(_capi.ANN_ERR_NEGATIVE_K, "numOfNeighbours < 0")
        if not self._h:  # nothing appended yet: BruteForceIndex.scala:76-89 yields an empty list
            return out_ids, out_dist, out_cnt
        _capi.check(_capi.lib().ann_query_batch(self._h, _ptr(q), b, q.shape[1], k, _ptr(out_ids), _ptr(out_dist),
                                                _ptr(out_cnt)))
        return out_ids, out_dist, out_cnt

    def query_batch_device(self, queries_t, k: int, out_ids_t, out_dist_t, out_count_t, stream: int = 0) -> None:
        """Device-pointer query: CUDA tensors in, CUDA tensors out, enqueued on `stream` without synchronising.
        Call raise_pending_error() after the stream has been synchronised."""
        self.flush()
        if True:
            if not self._h:   # nothing appended yet: an empty list per query (BruteForceIndex.scala:76-89), as on the host path
                out_ids_t.fill_(-1)
                out_dist_t.fill_(float("inf"))
                if out_count_t is not None:
                    out_count_t.zero_()
                return
            _capi.check(_capi.lib().ann_query_batch_device(
                self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
                ctypes.c_void_p(out_ids_t.data_ptr()), ctypes.c_void_p(out_dist_t.data_ptr()),
                None if out_count_t is None else ctypes.c_void_p(out_count_t.data_ptr()), ctypes.c_void_p(stream)))

    def query_seed_device(self, queries_t, k: int, seed_keys_t, stream: int = 0) -> None:
        """First half of a sharded query (`ann_query_seed_device`): prepare the batch, score a prefix of this shard and
        publish k bounds per query into `seed_keys_t` ([b, k] int32/uint32 CUDA tensor, normally peer-mapped memory)."""
        self.flush()
        if True:
            _capi.check(_capi.lib().ann_query_seed_device(
                self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
                ctypes.c_void_p(seed_keys_t.data_ptr()), ctypes.c_void_p(stream)))

    def query_finish_device(self, queries_t, k: int, peer_seed_key_ptrs, out_ids_t, out_dist_t, out_count_t, stream: int = 0) -> None:
        """Second half (`ann_query_finish_device`): `peer_seed_key_ptrs` are the device addresses of every shard's published
        key array as mapped into this process (own shard included); the shard is scored against the global threshold they
        imply and its candidates for the global top-k are written to the outputs (count may be < k)."""
        world = len(peer_seed_key_ptrs)
        arr = (ctypes.c_void_p * max(world, 1))(*[int(p) for p in peer_seed_key_ptrs])
        if True:
            _capi.check(_capi.lib().ann_query_finish_device(
                self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
                arr if world else None, world, ctypes.c_void_p(out_ids_t.data_ptr()), ctypes.c_void_p(out_dist_t.data_ptr()),
                None if out_count_t is None else ctypes.c_void_p(out_count_t.data_ptr()), ctypes.c_void_p(stream)))

    def query_seed_push_device(self, queries_t, k: int, dst_ptrs, stream: int = 0) -> None:
        """`ann_query_seed_push_device`: like `query_seed_device`, but the bounds are written into this shard's block of every
        peer's receive buffer (`dst_ptrs`: one device address per peer, own copy included)."""
        n = len(dst_ptrs)
        arr = (ctypes.c_void_p * n)(*[int(p) for p in dst_ptrs])
        self.flush()
        _capi.check(_capi.lib().ann_query_seed_push_device(
            self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k), arr, n,
            ctypes.c_void_p(stream)))

    def query_filter_push_device(self, queries_t, k: int, seed_src_ptrs, kth_dst_ptrs, stream: int = 0) -> None:
        """`ann_query_filter_push_device`: seed bounds read from the LOCAL receive buffer (`seed_src_ptrs`), the k best bounds
        pushed into every peer's receive buffer (`kth_dst_ptrs`)."""
        w, n = len(seed_src_ptrs), len(kth_dst_ptrs)
        src = (ctypes.c_void_p * max(w, 1))(*[int(p) for p in seed_src_ptrs])
        dst = (ctypes.c_void_p * n)(*[int(p) for p in kth_dst_ptrs])
        _capi.check(_capi.lib().ann_query_filter_push_device(
            self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
            src if w else None, w, dst, n, ctypes.c_void_p(stream)))

    def query_seed_slice_push_device(self, queries_t, k: int, q_begin: int, q_count: int, n_slices: int, dst_ptrs, stream: int = 0) -> None:
        """`ann_query_seed_slice_push_device` (sliced seeding): seed only the queries [q_begin, q_begin + q_count) over
        `n_slices` times the rows and write ONE bound per query into entries [q_begin, ...) of every peer's [b] bound array
        (`dst_ptrs`: the array's device address inside every peer, own copy included)."""
        n = len(dst_ptrs)
        arr = (ctypes.c_void_p * n)(*[int(p) for p in dst_ptrs])
        self.flush()
        _capi.check(_capi.lib().ann_query_seed_slice_push_device(
            self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k), int(q_begin),
            int(q_count), int(n_slices), arr, n, ctypes.c_void_p(stream)))

    def query_filter_bounds_push_device(self, queries_t, k: int, bounds_ptr: int, world: int, kth_dst_ptrs, stream: int = 0) -> None:
        """`ann_query_filter_bounds_push_device`: thresholds from the LOCAL [b] bound array the slice owners filled, then the
        chunks, the last compaction and the k best bounds pushed into every peer's receive buffer (`kth_dst_ptrs`)."""
        n = len(kth_dst_ptrs)
        dst = (ctypes.c_void_p * n)(*[int(p) for p in kth_dst_ptrs])
        _capi.check(_capi.lib().ann_query_filter_bounds_push_device(
            self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
            ctypes.c_void_p(int(bounds_ptr)), int(world), dst, n, ctypes.c_void_p(stream)))

    def query_filter_device(self, queries_t, k: int, peer_seed_key_ptrs, kth_keys_t, stream: int = 0) -> None:
        """Middle phase of the three-phase sharded query (`ann_query_filter_device`): global seed threshold, tensor-core
        chunks, last compaction; publishes this shard's k best bounds per query into `kth_keys_t` ([b, k] CUDA tensor)."""
        world = len(peer_seed_key_ptrs)
        arr = (ctypes.c_void_p * max(world, 1))(*[int(p) for p in peer_seed_key_ptrs])
        if True:
            _capi.check(_capi.lib().ann_query_filter_device(
                self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
                arr if world else None, world, ctypes.c_void_p(kth_keys_t.data_ptr()), ctypes.c_void_p(stream)))

    def query_rescore_device(self, queries_t, k: int, peer_kth_key_ptrs, out_ids_t, out_dist_t, out_count_t, stream: int = 0) -> None:
        """Last phase (`ann_query_rescore_device`): exact rescoring of this shard's rows under the k-th best bound of ALL
        shards; writes this shard's candidates for the global top-k (count may be < k; -1 = flagged, row invalid)."""
        world = len(peer_kth_key_ptrs)
        arr = (ctypes.c_void_p * max(world, 1))(*[int(p) for p in peer_kth_key_ptrs])
        if True:
            _capi.check(_capi.lib().ann_query_rescore_device(
                self._h, ctypes.c_void_p(queries_t.data_ptr()), int(queries_t.shape[0]), int(queries_t.shape[1]), int(k),
                arr if world else None, world, ctypes.c_void_p(out_ids_t.data_ptr()), ctypes.c_void_p(out_dist_t.data_ptr()),
                None if out_count_t is None else ctypes.c_void_p(out_count_t.data_ptr()), ctypes.c_void_p(stream)))

    def raise_pending_error(self) -> None:
        if not self._h:
            return
        v = ctypes.c_int64()
        _capi.check(_capi.lib().ann_get_stat(self._h, b"pending_error", ctypes.byref(v)))

    def query_with_distance(self, embedding, num_of_neighbors: int, runtime_params=BruteForceRuntimeParams) -> Future:
        def run():
            e = np.asarray(embedding, dtype=np.float32).reshape(1, -1)
            if num_of_neighbors <= 0:  # every push is popped again (BruteForceIndex.scala:83-85)
                return []
            ids, dist, cnt = self.batch_query_with_distance(e, num_of_neighbors)
            return [NeighborWithDistance(self.id_of(ids[0, j]), self.metric.from_absolute_distance(dist[0, j]))
                    for j in range(int(cnt[0]))]
        return self.future_pool(run)

    def query(self, embedding, num_of_neighbors: int, runtime_params=BruteForceRuntimeParams) -> Future:
        def run():
            e = np.asarray(embedding, dtype=np.float32).reshape(1, -1)
            if num_of_neighbors <= 0:
                return []
            ids, _, cnt = self.batch_query_with_distance(e, num_of_neighbors)
            return [self.id_of(ids[0, j]) for j in range(int(cnt[0]))]
        return self.future_pool(run)

    def read_rows(self, start: int, n: int):
        """Rows [start, start+n) and their device ids, in insertion order (what toDirectory iterates)."""
        self.flush()
        if True:
            ids = np.empty(n, dtype=np.int64)
            rows = np.empty((n, self.dim or 0), dtype=np.float32)
            if n:
                _capi.check(_capi.lib().ann_read_rows(self._h, start, n, _ptr(ids), _ptr(rows)))
            return ids, rows

    def loadtest(self, queries, k: int, threads: int, calls_per_thread: int, expect_ids=None) -> dict:
        """`ann_loadtest`: `threads` native host threads issue one-vector queries concurrently (the reference's load
        generator, service/loadtest/AnnLoadTestWorker.scala:92-115); returns rate and latency percentiles in microseconds."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        exp = None if expect_ids is None else np.ascontiguousarray(expect_ids, dtype=np.int64)
        st = _capi.AnnLoadStats()
        self.flush()
        _capi.check(_capi.lib().ann_loadtest(self._h, _ptr(q), q.shape[0], q.shape[1], int(k), int(threads), int(calls_per_thread),
                                             _ptr(exp), ctypes.byref(st)))
        return {f: getattr(st, f) for f, _ in _capi.AnnLoadStats._fields_}

    # ---- tuning / introspection ------------------------------------------------------------------------------
    def set_option(self, name: str, value: int) -> None:
        _capi.check(_capi.lib().ann_set_option(self._h, name.encode(), int(value)))

    def stat(self, name: str) -> int:
        v = ctypes.c_int64()
        _capi.check(_capi.lib().ann_get_stat(self._h, name.encode(), ctypes.byref(v)))
        return int(v.value)


def merge_topk_device(ids_t, dist_t, count_t, k: int, stream: int = 0):
    """K5 on CUDA tensors: ids/dist [S, b, k], counts [S, b] -> (ids [b,k], dist [b,k], counts [b]) on the same device.
    Replaces ComposedQueryable's flatten/sort/take (ShardApi.scala:77-85)."""
    import torch

    s, b = int(ids_t.shape[0]), int(ids_t.shape[1])
    kk = int(ids_t.shape[2])
    dev = ids_t.device
    out_ids = torch.full((b, kk), -1, dtype=torch.int64, device=dev)
    out_dist = torch.full((b, kk), float("inf"), dtype=torch.float32, device=dev)
    out_cnt = torch.zeros((b,), dtype=torch.int32, device=dev)
    if k > 0 and b > 0:
        assert kk == k
        _capi.check(_capi.lib().ann_merge_topk_device(
            dev.index or 0, ctypes.c_void_p(ids_t.data_ptr()), ctypes.c_void_p(dist_t.data_ptr()),
            ctypes.c_void_p(count_t.data_ptr()), s, b, k, ctypes.c_void_p(out_ids.data_ptr()),
            ctypes.c_void_p(out_dist.data_ptr()), ctypes.c_void_p(out_cnt.data_ptr()), ctypes.c_void_p(stream)))
    return out_ids, out_dist, out_cnt


# ------------------------------------------------------------------------------------------------ persistence
_MAGIC = b"B200ANN\x01"


class SerializableBruteForceIndex:
    """Mirror of SerializableBruteForceIndex / BruteForceDeserialization (BruteForceIndex.scala:94-162;
    BruteForceDeserialization.scala:18-64): one data file `BruteForceFileData` inside a directory, `_SUCCESS` marker
    (common/IndexOutputFile.scala:29,58-62).

    fmt="thrift" (default) is the reference's own format, written and read natively (`ann_save_directory` /
    `ann_load_directory`, csrc/persist.cu): back-to-back TBinaryProtocol `PersistedEmbedding{1: binary id, 2:
    embedding.Embedding}` structs with no header (ThriftIteratorIO.scala:14-22), ids as big-endian Long
    (AnnInjections.scala:8).  The inner `embedding.Embedding` thrift struct is not in the open-source tree; the assumed
    layout is stated in persist.cu and switchable with `layout` (the reader accepts every variant).
    fmt="raw" is the compact native layout kept from round 1 (little endian):
        magic "B200ANN\x01" | int32 metric ordinal | int32 dim | int64 n | int64 ids[n] | float32 rows[n][dim]
    `from_directory` recognises either by the magic.  Only native int64 ids are persisted."""

    DataFileName = BruteForceIndex.DataFileName
    SuccessMarker = "_SUCCESS"

    @staticmethod
    def to_directory(index: BruteForceIndex, directory, chunk_rows: int = 1 << 20, fmt: str = "thrift",
                     id_format: int = _capi.ANN_ID_INT64_BE, layout: int = _capi.ANN_LAYOUT_FLOAT_TENSOR) -> None:
        import os
        from pathlib import Path

        if not index.native_ids:
            raise TypeError("only native int64 ids can be persisted")
        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        n = index.size()
        if fmt == "thrift":
            if not index._h:
                (d / SerializableBruteForceIndex.DataFileName).write_bytes(b"")   # an empty stream is an empty index
                (d / SerializableBruteForceIndex.SuccessMarker).write_bytes(b"")
                return
            _capi.check(_capi.lib().ann_save_directory(index._h, os.fsencode(str(d)), id_format, layout))
            return
        if fmt != "raw":
            raise ValueError("fmt must be 'thrift' or 'raw'")
        dim = index.dim or 0
        with open(d / SerializableBruteForceIndex.DataFileName, "wb") as f:
            f.write(_MAGIC)
            f.write(np.array([index.metric.ordinal, dim], dtype="<i4").tobytes())
            f.write(np.array([n], dtype="<i8").tobytes())
            for s in range(0, n, chunk_rows):          # ids first, then rows, each streamed in chunks
                ids, _ = index.read_rows(s, min(chunk_rows, n - s))
                f.write(ids.astype("<i8").tobytes())
            for s in range(0, n, chunk_rows):
                _, rows = index.read_rows(s, min(chunk_rows, n - s))
                f.write(rows.astype("<f4").tobytes())
        (d / SerializableBruteForceIndex.SuccessMarker).write_bytes(b"")

    toDirectory = to_directory

    @staticmethod
    def from_directory(directory, metric: Metric, future_pool: FuturePool, *, device: int = 0,
                       chunk_rows: int = 1 << 20, dim: int = 0, id_format: int = _capi.ANN_ID_AUTO) -> BruteForceIndex:
        import os
        from pathlib import Path

        d = Path(directory)
        with open(d / SerializableBruteForceIndex.DataFileName, "rb") as f:
            head = f.read(8)
        if head != _MAGIC:      # the reference's thrift stream, decoded natively
            index = BruteForceIndex(metric, future_pool, device=device)
            cfg = _capi.AnnConfig(metric.ordinal, int(dim), 0, device, index._cfg["flags"])
            h = ctypes.c_void_p()
            _capi.check(_capi.lib().ann_load_directory(ctypes.byref(cfg), os.fsencode(str(d)), id_format, ctypes.byref(h)))
            index._h = h
            index.dim = index.stat("dim")
            index._n = index.size()
            index._version += 1
            return index
        with open(d / SerializableBruteForceIndex.DataFileName, "rb") as f:
            f.read(8)
            ordinal, dim_ = np.frombuffer(f.read(8), dtype="<i4")
            (n,) = np.frombuffer(f.read(8), dtype="<i8")
            if int(ordinal) != metric.ordinal:
                raise ValueError(f"index was written with metric ordinal {ordinal}, asked to load as {metric}")
            ids = np.frombuffer(f.read(int(n) * 8), dtype="<i8")
            index = BruteForceIndex(metric, future_pool, device=device, capacity_hint=int(n))
            for s in range(0, int(n), chunk_rows):
                m = min(chunk_rows, int(n) - s)
                rows = np.frombuffer(f.read(m * int(dim_) * 4), dtype="<f4").reshape(m, int(dim_))
                index.append_batch(ids[s:s + m], rows)
        return index

    fromDirectory = from_directory
