"""Generates tests/golden/*.npz: small seeded inputs together with the expected neighbour ids / distances.

PARITY UNPINNED: the reference has no golden vectors for ann/ and cannot run here (Scala/JVM, unshipped arithmetic), so
these fixtures are produced by the CPU restatement itself (oracle.c, cross-checked against oracle_np.py at generation
time) and pin the conventions C1-C7 against regressions -- they are NOT outputs of the reference.

    python oracle/gen_golden.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def case(name, metric, n, d, b, k, seed, special=False, dup=False, scale=1.0):
    rng = np.random.default_rng(seed)
    corpus = (rng.standard_normal((n, d)) * scale / np.sqrt(d)).astype(np.float32)
    if dup:
        corpus[n // 2: n // 2 + n // 10] = corpus[: n // 10]
    if special:
        corpus[1] = 0.0
        corpus[2, 0] = np.nan
        corpus[3, 1] = np.inf
        corpus[4, 2] = -np.inf
        corpus[5] = -0.0
    q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
    ids = (rng.permutation(n).astype(np.int64) - n // 3) * 1_000_003
    ci, cd, cc = oracle.query_canonical(metric, corpus, ids, q, k)
    ni, nd, nc = onp.query_canonical(metric, corpus, ids, q, k)
    assert (ci == ni).all() and (onp.float_order_key(cd) == onp.float_order_key(nd)).all() and (cc == nc).all(), name
    np.savez_compressed(OUT / f"{name}.npz", metric=np.int32(metric), corpus=corpus, ids=ids, queries=q, k=np.int32(k),
                        expect_ids=ci, expect_dist=cd, expect_count=cc)
    print(name, corpus.shape, q.shape, k)


def metric_case():
    """Metric.distance for plain pairs and MetricUtil.norm (Metric.scala:76-158, 285-289): one fixture for the direct
    entry points (ann_distance_pairs / ann_normalize_rows), special values included."""
    rng = np.random.default_rng(777)
    n, d = 96, 37
    a = (rng.standard_normal((n, d)) * 3).astype(np.float32)
    b = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    a[0], b[1] = 0.0, 0.0
    a[2, 0], b[3, -1] = np.nan, np.inf
    a[4] = b[4]
    a[5], b[5] = 1e-30, 1e-30
    a[6], b[6] = 3e18, 3e18
    a[7] = -0.0
    out = {"a": a, "b": b}
    for m, mn in ((oracle.L2, "l2"), (oracle.COSINE, "cosine"), (oracle.INNER_PRODUCT, "ip")):
        dist = np.array([oracle.distance(m, a[i], b[i]) for i in range(n)], np.float32)
        twin = np.array([onp.distances(m, a[i:i + 1], b[i])[0] for i in range(n)], np.float32)
        assert (onp.float_order_key(dist) == onp.float_order_key(twin)).all(), mn
        out[f"dist_{mn}"] = dist
    out["dist_l2_squared"] = np.array([oracle.distance(oracle.L2, a[i], b[i], l2_squared=1) for i in range(n)], np.float32)
    out["norm_a"] = oracle.normalize(a)
    assert (onp.float_order_key(out["norm_a"]) == onp.float_order_key(onp.normalize(a))).all()
    (OUT / "metric").mkdir(exist_ok=True)
    np.savez_compressed(OUT / "metric" / "pairs.npz", **out)
    print("metric/pairs", a.shape)


if __name__ == "__main__":
    metric_case()
    for m, mn in ((oracle.L2, "l2"), (oracle.COSINE, "cosine"), (oracle.INNER_PRODUCT, "ip")):
        case(f"{mn}_small", m, 300, 24, 6, 10, seed=100 + m)
        case(f"{mn}_k_gt_n", m, 40, 8, 3, 64, seed=200 + m)
        case(f"{mn}_dups", m, 400, 16, 5, 50, seed=300 + m, dup=True)
        case(f"{mn}_special", m, 120, 12, 4, 120, seed=400 + m, special=True)
        case(f"{mn}_d200", m, 600, 200, 8, 100, seed=500 + m)
        case(f"{mn}_tiny_scale", m, 256, 32, 4, 20, seed=600 + m, scale=1e-3)
