"""numpy twin of oracle.c.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

Restates the same reference lines independently (vectorised over rows, strictly sequential over the
dimension so that every fp64 rounding happens in the order oracle.c / the GPU rescoring kernel use):
  distances ............ ann/.../common/Metric.scala:88-94, 119-125, 150-158, 263-290
  total order .......... ann/.../common/Metric.scala:17-36 (Ordering.Float.compare)
  scan + top-k ......... ann/.../brute_force/BruteForceIndex.scala:66-91 (canonical refinement C5)
  shard merge .......... ann/.../common/ShardApi.scala:72-86
PARITY UNPINNED (no reference tests/fixtures for ann/; arithmetic unshipped) -- conventions C1..C7.
"""
from __future__ import annotations

import numpy as np

L2, COSINE, INNER_PRODUCT = 0, 1, 2


def float_order_key(x) -> np.ndarray:
    """java.lang.Float.compare as a uint32 key: -0.0 < +0.0 < ... < +inf < NaN (all NaN equal)."""
    x = np.asarray(x, dtype=np.float32)
    b = x.view(np.uint32)
    key = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)
    return np.where(np.isnan(x), np.uint32(0xFFFFFFFF), key)


def _seq_sum_f64(terms_fn, n_rows: int, d: int) -> np.ndarray:
    acc = np.zeros(n_rows, dtype=np.float64)
    for i in range(d):
        acc = acc + terms_fn(i)
    return acc


def distances(metric: int, corpus: np.ndarray, query: np.ndarray, l2_squared: bool = False) -> np.ndarray:
    """distance(row, query) for every row, fp64 sequential accumulation, one rounding to fp32 (C1-C4)."""
    a = np.ascontiguousarray(corpus, dtype=np.float32)
    q = np.ascontiguousarray(query, dtype=np.float32)
    n, d = a.shape
    a64 = a.astype(np.float64)
    q64 = q.astype(np.float64)
    with np.errstate(all="ignore"):
        if metric == L2:
            def term(i):
                diff = a64[:, i] - q64[i]
                return diff * diff
            acc = _seq_sum_f64(term, n, d)
            return (acc if l2_squared else np.sqrt(acc)).astype(np.float32)
        dot = _seq_sum_f64(lambda i: a64[:, i] * q64[i], n, d)
        if metric == INNER_PRODUCT:
            return (np.float32(1.0) - dot.astype(np.float32)).astype(np.float32)
        na = _seq_sum_f64(lambda i: a64[:, i] * a64[:, i], n, d)
        nb = np.float64(0.0)
        for i in range(d):
            nb = nb + q64[i] * q64[i]
        cs = dot / (np.sqrt(na) * np.sqrt(nb))
        return (np.float32(1.0) - cs.astype(np.float32)).astype(np.float32)


def normalize(rows) -> np.ndarray:
    """MetricUtil.norm (Metric.scala:285-289), convention C8: fp64 sequential squared norm, fp64 divide, one rounding."""
    a = np.atleast_2d(np.ascontiguousarray(rows, dtype=np.float32))
    a64 = a.astype(np.float64)
    n2 = _seq_sum_f64(lambda i: a64[:, i] * a64[:, i], a.shape[0], a.shape[1])
    with np.errstate(all="ignore"):
        return (a64 / np.sqrt(n2)[:, None]).astype(np.float32)


def query_canonical(metric: int, corpus, ids, queries, k: int, l2_squared: bool = False):
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    queries = np.atleast_2d(np.ascontiguousarray(queries, dtype=np.float32))
    n = corpus.shape[0]
    ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    b = queries.shape[0]
    kk = max(k, 0)
    out_ids = np.full((b, kk), -1, dtype=np.int64)
    out_dist = np.full((b, kk), np.inf, dtype=np.float32)
    out_cnt = np.zeros(b, dtype=np.int32)
    for qi in range(b):
        dist = distances(metric, corpus, queries[qi], l2_squared) if n else np.zeros(0, np.float32)
        order = np.lexsort((ids, float_order_key(dist)))[:kk]
        m = order.shape[0]
        out_ids[qi, :m] = ids[order]
        out_dist[qi, :m] = dist[order]
        out_cnt[qi] = m
    return out_ids, out_dist, out_cnt


def merge_canonical(in_ids, in_dist, in_count, k: int):
    ids = np.concatenate([np.asarray(in_ids[s][: in_count[s]], dtype=np.int64) for s in range(len(in_count))])
    dist = np.concatenate([np.asarray(in_dist[s][: in_count[s]], dtype=np.float32) for s in range(len(in_count))])
    order = np.lexsort((ids, float_order_key(dist)))[: max(k, 0)]
    return ids[order], dist[order]
