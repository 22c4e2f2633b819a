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
