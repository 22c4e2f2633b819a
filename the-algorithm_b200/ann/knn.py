"""Batched offline kNN driver over the GPU index -- the caller that actually presents query batches in the thousands.

Mirrors, for one box, what the reference does on Hadoop:
  find_nearest_neighbours ......... KnnHelper.findNearestNeighbours / findNearestNeighboursWithIndexingStrategy
                                    (ann/src/main/scala/com/twitter/ann/scalding/offline/KnnHelper.scala:168-215, 248-347):
                                    build an index over the search space, query every query embedding, keep k per query
  nearest_neighbors_to_string ..... KnnHelper.nearestNeighborsToString (:415-429):
                                    "queryId<TAB>neighborId:distance<TAB>neighborId:distance..." ascending by distance
  write_truth_set / load_truth_set  KnnTruthSetGenerator (KnnTruthSetGenerator.scala:58-70) and
                                    LoadTestUtils.getTruthSetMap (service/loadtest/LoadTestUtils.scala:42-58)
  recall .......................... LoadTestRecorder's recall@n (service/loadtest/LoadTestRecorder.scala:22-61):
                                    |truth[:n] ∩ result| / n

The reference shards the search space into random "search groups" across reducers and merges with sortedTake
(KnnHelper.scala:272-346); here the search space lives on the GPU(s) and queries stream through in tiles, so the merge across
corpus tiles is only needed when the corpus exceeds one shard (then it is the same (distance, id) merge as everywhere else).
All arithmetic happens in the CUDA engine; this module only tiles, formats and counts.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .brute_force import BruteForceIndex
from .common import FuturePool, Metric


def java_float_to_string(x) -> str:
    """java.lang.Float.toString for the `${distance.distance}` interpolation in KnnHelper.scala:425: decimal notation for
    1e-3 <= |x| < 1e7 with at least one fractional digit, otherwise computerised scientific notation (d.dddE-n), the
    shortest digits that round-trip the float (JDK < 19 occasionally prints one digit more, e.g. Float.MIN_NORMAL;
    readers of the truth set only parse the ids, LoadTestUtils.scala:52)."""
    f = np.float32(x)
    if np.isnan(f):
        return "NaN"
    if np.isinf(f):
        return "Infinity" if f > 0 else "-Infinity"
    if f == 0:
        return "-0.0" if math.copysign(1.0, float(f)) < 0 else "0.0"
    a = abs(float(f))
    if 1e-3 <= a < 1e7:
        s = np.format_float_positional(f, unique=True, trim="0")
        return s if "." in s else s + ".0"
    s = np.format_float_scientific(f, unique=True, trim="0", exp_digits=1)   # e.g. 1.e-04 / 1.5e+08
    mant, exp = s.split("e")
    if mant.endswith("."):
        mant += "0"
    return f"{mant}E{int(exp)}"


def nearest_neighbors_to_string(query_id, neighbors: Sequence[Tuple[object, float]], id_distance_separator: str = ":",
                                neighbor_separator: str = "\t") -> str:
    """KnnHelper.nearestNeighborsToString (KnnHelper.scala:415-429)."""
    parts = [str(query_id)] + [f"{nid}{id_distance_separator}{java_float_to_string(d)}" for nid, d in neighbors]
    return neighbor_separator.join(parts)


def find_nearest_neighbours(query_ids: Sequence, query_embeddings, index_ids, index_embeddings, metric: Metric,
                            num_neighbors: int, *, device: int = 0, query_tile: int = 4096, index: Optional[BruteForceIndex] = None
                            ) -> Iterator[Tuple[object, List[Tuple[int, float]]]]:
    """Yields (queryId, [(neighborId, distance) ...]) for every query, nearest first -- the TypedPipe the reference's
    findNearestNeighbours returns (KnnHelper.scala:168-215), computed with whole query tiles per device call."""
    q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
    own = index is None
    if own:
        index = BruteForceIndex.apply(metric, FuturePool.immediate_pool(), device=device,
                                      capacity_hint=int(np.asarray(index_embeddings).shape[0]))
        index.append_batch(index_ids, index_embeddings)
    try:
        for t0 in range(0, q.shape[0], query_tile):
            ids, dist, cnt = index.batch_query_with_distance(q[t0:t0 + query_tile], num_neighbors)
            for j in range(ids.shape[0]):
                c = int(cnt[j])
                yield query_ids[t0 + j], [(index.id_of(ids[j, i]), float(dist[j, i])) for i in range(c)]
    finally:
        if own:
            index.close()


def knn_join(query_embeddings, index_ids, index_embeddings, metric: Metric, num_neighbors: int, *, device: int = 0,
             query_tile: int = 4096, corpus_tile_rows: int = 0, l2_squared: bool = False):
    """`ann_knn_join`: the whole job in ONE native call -- corpus tiles that fit the device, query tiles streaming through
    with copies overlapped with the kernels, per-corpus-tile lists merged on the device.  Returns (ids [nq,k] int64,
    distances [nq,k] float32, counts [nq] int32) with `index_ids` as given (int64).  `find_nearest_neighbours` above is the
    same computation driven tile by tile from Python through an index object."""
    import ctypes

    from .. import _capi

    q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
    rows = np.ascontiguousarray(index_embeddings, dtype=np.float32)
    ids = np.ascontiguousarray(index_ids, dtype=np.int64)
    if q.ndim != 2 or rows.ndim != 2 or (rows.shape[0] and rows.shape[1] != q.shape[1]) or ids.shape != (rows.shape[0],):
        raise ValueError("knn_join: queries [nq, d], rows [n, d], ids [n] expected")
    nq, k = q.shape[0], int(num_neighbors)
    out_ids = np.empty((nq, max(k, 0)), np.int64)
    out_dist = np.empty((nq, max(k, 0)), np.float32)
    out_cnt = np.zeros((nq,), np.int32)
    cfg = _capi.AnnConfig(metric.ordinal, q.shape[1], 0, device, _capi.ANN_FLAG_L2_SQUARED if l2_squared else 0)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)   # noqa: E731
    _capi.check(_capi.lib().ann_knn_join(ctypes.byref(cfg), p(ids), p(rows), rows.shape[0], p(q), nq, k, corpus_tile_rows,
                                         query_tile, p(out_ids), p(out_dist), p(out_cnt)))
    return out_ids, out_dist, out_cnt


def write_truth_set(path, results: Iterable[Tuple[object, Sequence[Tuple[object, float]]]]) -> int:
    """One TSV line per query, KnnTruthSetGenerator's output (`TypedText.tsv(knnOutputPath)`, :58-70).  Returns the line count."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    n = 0
    with open(path, "w", encoding="utf-8") as f:
        for qid, neighbors in results:
            f.write(nearest_neighbors_to_string(qid, neighbors) + "\n")
            n += 1
    return n


def load_truth_set(path, query_converter=int, index_converter=int) -> Dict[object, List[object]]:
    """LoadTestUtils.getTruthSetMap (:42-58): id -> neighbour ids; the distance after the last ':' is dropped."""
    out: Dict[object, List[object]] = {}
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            arr = line.rstrip("\n").split("\t")
            if not arr or not arr[0]:
                continue
            out[query_converter(arr[0])] = [index_converter(s[: s.rindex(":")]) for s in arr[1:]]
    assert out, f"Must have some something in the truth set {path}"
    return out


def recall(truth: Sequence, result: Sequence, top_n: Optional[int] = None) -> float:
    """LoadTestRecorder recall (:22-61): fraction of the first `top_n` true neighbours present in the result."""
    t = list(truth)[: top_n or len(truth)]
    if not t:
        return 1.0
    r = set(result)
    return sum(1 for x in t if x in r) / len(t)
