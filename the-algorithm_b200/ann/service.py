"""Query-service shell pieces around a GPU Queryable.

The reference serves ANN queries one vector per RPC: QueryIndexThriftController.query decodes a NearestNeighborQuery and calls
queryable.queryWithDistance(embedding, k, params) (ann/src/main/scala/com/twitter/ann/service/query_server/common/
QueryIndexThriftController.scala:39-90).  A GPU index wants batches, so the piece that matters is a micro-batcher behind the
unchanged single-vector `Queryable` trait: concurrent callers are coalesced into one device call.

  MicroBatchingQueryable .... Queryable facade: query / query_with_distance return Futures that complete when the batch the
                              request joined has been answered by `batch_query_with_distance`
  warmup .................... the server warm-up contract: random uniform [-1, 1) queries with k = 100 until 100 successes
                              (ann/.../query_server/hnsw/HnswQueryIndexServer.scala:85-97; common/warmup/Warmup.scala:45-47)
The thrift transport itself (ann_common.thrift:118-169) is out of scope: this is the in-process half of the server.
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import Future
from typing import List, Optional

import numpy as np

from .common import NeighborWithDistance, Queryable


class MicroBatchingQueryable(Queryable):
    def __init__(self, queryable, max_batch: int = 256, max_delay_ms: float = 1.0):
        if not hasattr(queryable, "batch_query_with_distance"):
            raise TypeError("the wrapped Queryable needs batch_query_with_distance")
        self.inner = queryable
        self.metric = queryable.metric
        self.max_batch = int(max_batch)
        self.max_delay = max_delay_ms / 1e3
        self._cv = threading.Condition()
        self._queue: List = []          # (embedding, k, want_distance, future)
        self._closed = False
        self.batches = 0
        self.requests = 0
        self._t = threading.Thread(target=self._run, name="b200ann-microbatch", daemon=True)
        self._t.start()

    # ---- Queryable ----------------------------------------------------------------------------------------------
    def _submit(self, embedding, k: int, want_distance: bool) -> Future:
        f: Future = Future()
        if k <= 0:  # BruteForceIndex.scala:83-85: every push is popped again
            f.set_result([])
            return f
        e = np.asarray(embedding, dtype=np.float32).reshape(-1)
        with self._cv:
            if self._closed:
                f.set_exception(RuntimeError("MicroBatchingQueryable is closed"))
                return f
            self._queue.append((e, int(k), want_distance, f))
            self._cv.notify()
        return f

    def query_with_distance(self, embedding, num_of_neighbors: int, runtime_params=None) -> Future:
        return self._submit(embedding, num_of_neighbors, True)

    def query(self, embedding, num_of_neighbors: int, runtime_params=None) -> Future:
        return self._submit(embedding, num_of_neighbors, False)

    # ---- batching loop ------------------------------------------------------------------------------------------
    def _run(self):
        while True:
            with self._cv:
                while not self._queue and not self._closed:
                    self._cv.wait()
                if self._closed and not self._queue:
                    return
                deadline = time.monotonic() + self.max_delay
                while len(self._queue) < self.max_batch and not self._closed:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                batch, self._queue = self._queue[: self.max_batch], self._queue[self.max_batch:]
            self._answer(batch)

    def _answer(self, batch):
        self.batches += 1
        self.requests += len(batch)
        by_k = {}
        for item in batch:
            by_k.setdefault((item[1], item[0].shape[0]), []).append(item)
        for (k, _dim), items in by_k.items():
            try:
                ids, dist, cnt = self.inner.batch_query_with_distance(np.stack([it[0] for it in items]), k)
                for j, (_, _, want_distance, fut) in enumerate(items):
                    c = int(cnt[j])
                    if want_distance:
                        fut.set_result([NeighborWithDistance(self.inner.id_of(ids[j, i]), self.metric.from_absolute_distance(dist[j, i]))
                                        for i in range(c)])
                    else:
                        fut.set_result([self.inner.id_of(ids[j, i]) for i in range(c)])
            except BaseException as e:  # a failed device call fails every request of that group, like a failed Future
                for it in items:
                    if not it[3].done():
                        it[3].set_exception(e)

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._t.join(timeout=10)

    @property
    def mean_batch_size(self) -> float:
        return self.requests / self.batches if self.batches else 0.0


def warmup(queryable: Queryable, dimension: int, k: int = 100, successes: int = 100, timeout_ms: float = 50.0,
           max_attempts: int = 10_000, seed: Optional[int] = 0) -> int:
    """Warmup.run (common/warmup/Warmup.scala:15-50) with HnswQueryIndexServer's parameters (:85-97): random uniform
    [-1, 1) vectors, k = 100, until `successes` queries returned within `timeout_ms`.  Returns the attempts used."""
    rng = np.random.default_rng(seed)
    ok = attempts = 0
    while ok < successes and attempts < max_attempts:
        attempts += 1
        q = rng.uniform(-1.0, 1.0, dimension).astype(np.float32)
        t0 = time.monotonic()
        try:
            queryable.query_with_distance(q, k, None).result(timeout=timeout_ms / 1e3 * 20)
            if (time.monotonic() - t0) * 1e3 <= timeout_ms:
                ok += 1
        except Exception:
            pass
    return attempts
