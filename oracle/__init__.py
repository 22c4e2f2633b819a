"""CPU oracle for the exact-kNN path.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this package.  PARITY UNPINNED: the reference has no tests or fixtures for ann/ and its
arithmetic is unshipped; the conventions are C1..C7 of SURVEY.md section 8(c).

Two restatements live here and are checked against each other in tests/test_oracle.py:
  * ``oracle.c``      -- C, built by ``oracle/Makefile`` into ``oracle/_build/liboracle.so``
  * ``oracle_np.py``  -- numpy twin (vectorised over rows, sequential over the dimension)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

L2, COSINE, INNER_PRODUCT = 0, 1, 2
METRIC_BY_NAME = {"L2": L2, "Cosine": COSINE, "InnerProduct": INNER_PRODUCT}


_FAST_PATH = _HERE / "_build" / "libfastcpu.so"


def build(force: bool = False) -> Path:
    stale = any(not lib_.exists() or lib_.stat().st_mtime < (_HERE / src).stat().st_mtime
                for lib_, src in ((_LIB_PATH, "oracle.c"), (_FAST_PATH, "fast_cpu.c")))
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-s", "all"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists() or os.environ.get("ORACLE_REBUILD"):
            build(force=bool(os.environ.get("ORACLE_REBUILD")))
        elif _LIB_PATH.stat().st_mtime < (_HERE / "oracle.c").stat().st_mtime:
            try:
                build()
            except Exception:  # GPU box without make: use the shipped build
                pass
        L = ctypes.CDLL(str(_LIB_PATH))
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        cp = ctypes.POINTER(ctypes.c_int32)
        L.oracle_distance.restype = ctypes.c_float
        L.oracle_distance.argtypes = [ctypes.c_int, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_normalize.restype = None
        L.oracle_normalize.argtypes = [fp, ctypes.c_int, fp]
        L.oracle_float_order_key.restype = ctypes.c_uint32
        L.oracle_float_order_key.argtypes = [ctypes.c_float]
        L.oracle_query_canonical.restype = ctypes.c_int
        L.oracle_query_canonical.argtypes = [ctypes.c_int, fp, ip, ctypes.c_int64, ctypes.c_int, fp, ctypes.c_int,
                                             ctypes.c_int, ip, fp, cp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_query_fast_cpu.restype = ctypes.c_int
        L.oracle_query_fast_cpu.argtypes = [ctypes.c_int, fp, ip, ctypes.c_int64, ctypes.c_int, fp, ctypes.c_int,
                                            ctypes.c_int, ip, fp, ctypes.c_int]
        L.oracle_index_create.restype = ctypes.c_void_p
        L.oracle_index_create.argtypes = [ctypes.c_int, ctypes.c_int]
        L.oracle_index_append.restype = ctypes.c_int
        L.oracle_index_append.argtypes = [ctypes.c_void_p, ip, fp, ctypes.c_int64]
        L.oracle_index_size.restype = ctypes.c_int64
        L.oracle_index_size.argtypes = [ctypes.c_void_p]
        L.oracle_index_destroy.restype = None
        L.oracle_index_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_index_query.restype = ctypes.c_int
        L.oracle_index_query.argtypes = [ctypes.c_void_p, fp, ctypes.c_int, ctypes.c_int, ip, fp, cp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int]
        L.oracle_merge.restype = ctypes.c_int
        L.oracle_merge.argtypes = [ip, fp, cp, ctypes.c_int, ctypes.c_int, ip, fp, cp, ctypes.c_int]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def distance(metric: int, row, query, accum: int = 0, l2_squared: int = 0) -> np.float32:
    r, rp = _f(row)
    q, qp = _f(query)
    return np.float32(lib().oracle_distance(metric, rp, qp, r.shape[-1], accum, l2_squared))


def normalize(rows) -> np.ndarray:
    """MetricUtil.norm (Metric.scala:285-289) under convention C8, row by row."""
    r, _ = _f(np.atleast_2d(rows))
    out = np.empty_like(r)
    fp = ctypes.POINTER(ctypes.c_float)
    for i in range(r.shape[0]):
        lib().oracle_normalize(r[i].ctypes.data_as(fp), r.shape[1], out[i].ctypes.data_as(fp))
    return out


def query_canonical(metric: int, corpus, ids, queries, k: int, accum: int = 0, l2_squared: int = 0,
                    nthreads: int = 0):
    """Exact top-k under the canonical (Float.compare(distance), id) order.  Returns (ids, dist, count)."""
    c, cp_ = _f(corpus)
    q, qp = _f(np.atleast_2d(queries))
    n, d = c.shape if c.ndim == 2 else (0, q.shape[1])
    b = q.shape[0]
    kk = max(k, 0)
    if ids is None:
        ids = np.arange(n, dtype=np.int64)
    i, ip_ = _i(ids)
    out_ids = np.full((b, kk), -1, dtype=np.int64)
    out_dist = np.full((b, kk), np.inf, dtype=np.float32)
    out_cnt = np.zeros(b, dtype=np.int32)
    lib().oracle_query_canonical(metric, cp_, ip_, n, d, qp, b, k,
                                 out_ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                 out_dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                 out_cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), accum, l2_squared, nthreads)
    return out_ids, out_dist, out_cnt


_fast = None


def query_blocked_cpu(metric: int, corpus, ids, queries, k: int, nthreads: int = 0):
    """The "fair CPU" figure (oracle/fast_cpu.c): query-blocked, row-parallel, SIMD, -O3 -ffast-math.  TIMED ONLY: fp32
    lane-parallel sums, so ids can differ from the oracle where distances are within fp32 noise of each other."""
    global _fast
    if _fast is None:
        if not _FAST_PATH.exists():
            build()
        _fast = ctypes.CDLL(str(_FAST_PATH))
        _fast.fastcpu_query.restype = ctypes.c_int
        _fast.fastcpu_query.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64), ctypes.c_int64,
                                        ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    c, cp_ = _f(corpus)
    q, qp = _f(np.atleast_2d(queries))
    n, d = c.shape
    b = q.shape[0]
    ip_ = None
    if ids is not None:
        i, ip_ = _i(ids)
    out_ids = np.full((b, k), -1, dtype=np.int64)
    out_dist = np.full((b, k), np.inf, dtype=np.float32)
    _fast.fastcpu_query(metric, cp_, ip_, n, d, qp, b, k, out_ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                        out_dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nthreads)
    return out_ids, out_dist


def query_fast_cpu(metric: int, corpus, ids, queries, k: int, nthreads: int = 0):
    """Contiguous multi-threaded fp32 CPU scan; timed only (the 'fair CPU' figure), never a parity oracle."""
    c, cp_ = _f(corpus)
    q, qp = _f(np.atleast_2d(queries))
    n, d = c.shape
    b = q.shape[0]
    if ids is None:
        ids = np.arange(n, dtype=np.int64)
    i, ip_ = _i(ids)
    out_ids = np.full((b, k), -1, dtype=np.int64)
    out_dist = np.full((b, k), np.inf, dtype=np.float32)
    lib().oracle_query_fast_cpu(metric, cp_, ip_, n, d, qp, b, k,
                                out_ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                out_dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nthreads)
    return out_ids, out_dist


class FaithfulIndex:
    """Reference-faithful BruteForceIndex restatement: linked list of heap-allocated rows + Scala 2.12
    PriorityQueue mechanics (BruteForceIndex.scala:26-92).  Ties resolve by heap history, not by id."""

    def __init__(self, metric: int, dim: int):
        self.metric, self.dim = metric, dim
        self._h = lib().oracle_index_create(metric, dim)

    def append(self, ids, rows):
        r, rp = _f(np.atleast_2d(rows))
        assert r.shape[1] == self.dim
        i, ip_ = _i(ids)
        lib().oracle_index_append(self._h, ip_, rp, r.shape[0])

    def size(self) -> int:
        return int(lib().oracle_index_size(self._h))

    def query(self, queries, k: int, accum: int = 0, l2_squared: int = 0, nthreads: int = 0):
        q, qp = _f(np.atleast_2d(queries))
        b = q.shape[0]
        kk = max(k, 0)
        out_ids = np.full((b, kk), -1, dtype=np.int64)
        out_dist = np.full((b, kk), np.inf, dtype=np.float32)
        out_cnt = np.zeros(b, dtype=np.int32)
        lib().oracle_index_query(self._h, qp, b, k, out_ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                 out_dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                 out_cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), accum, l2_squared, nthreads)
        return out_ids, out_dist, out_cnt

    def close(self):
        if self._h:
            lib().oracle_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge(in_ids, in_dist, in_count, k: int, faithful: bool = False):
    """ComposedQueryable merge of S shard lists (ShardApi.scala:77-85).  in_*: [S, k]."""
    i, ip_ = _i(in_ids)
    d, dp = _f(in_dist)
    c = np.ascontiguousarray(in_count, dtype=np.int32)
    s = i.shape[0]
    kk = max(k, 0)
    out_ids = np.full(kk, -1, dtype=np.int64)
    out_dist = np.full(kk, np.inf, dtype=np.float32)
    out_cnt = np.zeros(1, dtype=np.int32)
    lib().oracle_merge(ip_, dp, c.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), s, k,
                       out_ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                       out_dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                       out_cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(faithful))
    return out_ids, out_dist, int(out_cnt[0])


def max_threads() -> int:
    return int(lib().oracle_max_threads())
