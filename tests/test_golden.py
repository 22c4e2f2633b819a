"""The oracle against the committed fixtures (CPU), and the CUDA path against the same fixtures (GPU).
Fixtures are self-generated (oracle/gen_golden.py): the reference has none for ann/ -- parity unpinned."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

GOLDEN = sorted((Path(__file__).parent / "golden").glob("*.npz"))


def key(x):
    return onp.float_order_key(x)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_matches_golden(path):
    z = np.load(path)
    i, d, c = oracle.query_canonical(int(z["metric"]), z["corpus"], z["ids"], z["queries"], int(z["k"]))
    assert (i == z["expect_ids"]).all() and (key(d) == key(z["expect_dist"])).all() and (c == z["expect_count"]).all()


@pytest.mark.parametrize("path", [p for p in GOLDEN if "d200" not in p.stem], ids=lambda p: p.stem)
def test_numpy_twin_matches_golden(path):
    z = np.load(path)
    i, d, c = onp.query_canonical(int(z["metric"]), z["corpus"], z["ids"], z["queries"], int(z["k"]))
    assert (i == z["expect_ids"]).all() and (key(d) == key(z["expect_dist"])).all() and (c == z["expect_count"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_gpu_matches_golden(path):
    from the_algorithm_b200.ann.brute_force import BruteForceIndex
    from the_algorithm_b200.ann.common import FuturePool, Metric

    z = np.load(path)
    ix = BruteForceIndex.apply(Metric.from_thrift(int(z["metric"])), FuturePool.immediate_pool())
    ix.append_batch(z["ids"], z["corpus"])
    i, d, c = ix.batch_query_with_distance(z["queries"], int(z["k"]))
    ix.close()
    assert (c == z["expect_count"]).all()
    assert (i == z["expect_ids"]).all()
    assert (key(d) == key(z["expect_dist"])).all()


# ---- Metric.distance / MetricUtil.norm for plain vectors (tests/golden/metric/pairs.npz) ----
PAIRS = Path(__file__).parent / "golden" / "metric" / "pairs.npz"
METRICS = ((oracle.L2, "l2"), (oracle.COSINE, "cosine"), (oracle.INNER_PRODUCT, "ip"))


def test_oracle_metric_matches_golden_pairs():
    z = np.load(PAIRS)
    a, b = z["a"], z["b"]
    for m, mn in METRICS:
        got = np.array([oracle.distance(m, a[i], b[i]) for i in range(len(a))], np.float32)
        twin = np.array([onp.distances(m, a[i:i + 1], b[i])[0] for i in range(len(a))], np.float32)
        assert (key(got) == key(z[f"dist_{mn}"])).all() and (key(twin) == key(z[f"dist_{mn}"])).all()
    sq = np.array([oracle.distance(oracle.L2, a[i], b[i], l2_squared=1) for i in range(len(a))], np.float32)
    assert (key(sq) == key(z["dist_l2_squared"])).all()
    assert (key(oracle.normalize(a)) == key(z["norm_a"])).all() and (key(onp.normalize(a)) == key(z["norm_a"])).all()


@pytest.mark.gpu
def test_gpu_metric_matches_golden_pairs():
    from the_algorithm_b200.ann.common import Metric, MetricUtil

    z = np.load(PAIRS)
    for m, mn in METRICS:
        assert (key(Metric.from_thrift(m).distances(z["a"], z["b"])) == key(z[f"dist_{mn}"])).all()
    assert (key(Metric.from_thrift(oracle.L2).distances(z["a"], z["b"], l2_squared=True)) == key(z["dist_l2_squared"])).all()
    assert (key(MetricUtil.norm(z["a"])) == key(z["norm_a"])).all()
