"""§8f rows: the offline kNN driver (truth-set TSV in the reference's format) and the raw on-disk format."""
import math

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp
from the_algorithm_b200.ann.knn import (java_float_to_string, load_truth_set, nearest_neighbors_to_string, recall,
                                        write_truth_set)


def test_java_float_to_string_matches_float_tostring_rules():
    cases = {1.0: "1.0", 0.5: "0.5", -31.0: "-31.0", 100.0: "100.0", 1234567.0: "1234567.0", 1.0e7: "1.0E7", 1.5e8: "1.5E8",
             0.001: "0.001", 1.0e-4: "1.0E-4", 9.999e-4: "9.999E-4", 3.4028235e38: "3.4028235E38", 0.1: "0.1",
             np.float32(1.0) - np.float32(1 / math.sqrt(2)): "0.29289323", 1.1754944e-38: "1.1754944E-38"}
    for x, want in cases.items():
        assert java_float_to_string(x) == want, (x, java_float_to_string(x), want)
    assert java_float_to_string(float("nan")) == "NaN" and java_float_to_string(float("-inf")) == "-Infinity"
    assert java_float_to_string(0.0) == "0.0" and java_float_to_string(-0.0) == "-0.0"
    for v in np.random.default_rng(0).standard_normal(200).astype(np.float32) * 1e3:   # shortest digits round-trip
        assert np.float32(java_float_to_string(v).replace("E", "e")) == v


def test_truth_set_tsv_round_trip(tmp_path):
    res = [(7, [(3, 0.25), (9, 1.0e-4)]), (8, []), (11, [(5, -2.0)])]
    line = nearest_neighbors_to_string(*res[0])
    assert line == "7\t3:0.25\t9:1.0E-4"                       # KnnHelper.scala:415-429
    p = tmp_path / "knn" / "part-00000.tsv"
    assert write_truth_set(p, res) == 3
    m = load_truth_set(p)                                        # LoadTestUtils.scala:42-58: ids only
    assert m == {7: [3, 9], 8: [], 11: [5]}
    assert recall(m[7], [9, 4, 3]) == 1.0 and recall(m[7], [9], top_n=1) == 0.0 and recall([1, 2, 3, 4], [2, 4]) == 0.5


@pytest.mark.gpu
def test_find_nearest_neighbours_matches_oracle_and_round_trips(tmp_path):
    from the_algorithm_b200.ann.common import Cosine
    from the_algorithm_b200.ann.knn import find_nearest_neighbours

    rng = np.random.default_rng(3)
    corpus = (rng.standard_normal((30_000, 64)) / 8).astype(np.float32)
    ids = rng.permutation(30_000).astype(np.int64) + 10 ** 9
    q = rng.uniform(-1, 1, (700, 64)).astype(np.float32)
    qids = np.arange(700) * 3
    res = list(find_nearest_neighbours(qids, q, ids, corpus, Cosine, 20, query_tile=256))
    oi, od, _ = oracle.query_canonical(oracle.COSINE, corpus, ids, q, 20)
    assert [r[0] for r in res] == qids.tolist()
    assert [[n for n, _ in r[1]] for r in res] == oi.tolist()
    assert [[np.float32(d) for _, d in r[1]] for r in res] == od.tolist()
    p = tmp_path / "truth.tsv"
    write_truth_set(p, res)
    truth = load_truth_set(p)
    assert all(truth[int(qids[j])] == oi[j].tolist() for j in range(700))
    assert recall(truth[0], oi[0].tolist(), 10) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("metric_name", ["L2", "Cosine", "InnerProduct"])
@pytest.mark.parametrize("corpus_tile,query_tile", [(0, 4096), (7000, 256), (2999, 100), (30_000, 33)])
def test_native_knn_join_matches_oracle(metric_name, corpus_tile, query_tile):
    """ann_knn_join (KnnHelper.findNearestNeighbours in one native call): one or several corpus tiles merged on the
    device, ragged query tiles through the double-buffered pipeline -- always the single-index answer, bit for bit."""
    from the_algorithm_b200.ann.common import Metric
    from the_algorithm_b200.ann.knn import knn_join

    metric = Metric.from_string(metric_name)
    rng = np.random.default_rng(8)
    corpus = (rng.standard_normal((20_001, 48)) / 7).astype(np.float32)
    corpus[15_000:15_040] = corpus[:40]                      # exact ties across corpus tiles
    ids = rng.permutation(20_001).astype(np.int64) * 7 - 3
    q = rng.uniform(-1, 1, (777, 48)).astype(np.float32)
    gi, gd, gc = knn_join(q, ids, corpus, metric, 25, query_tile=query_tile, corpus_tile_rows=corpus_tile)
    oi, od, oc = oracle.query_canonical(metric.ordinal, corpus, ids, q, 25)
    assert (gc == oc).all() and (gi == oi).all()
    assert (gd.view(np.uint32) == od.view(np.uint32)).all()


@pytest.mark.gpu
def test_native_knn_join_edge_cases_and_flagged_queries():
    """k > rows of a corpus tile, k = 0, empty corpus, and queries the bounded selector cannot answer (a NaN query; a
    corpus of identical rows where everything ties): those tiles are redone through the exact fallback."""
    from the_algorithm_b200 import _capi
    from the_algorithm_b200.ann.common import InnerProduct, L2
    from the_algorithm_b200.ann.knn import knn_join

    rng = np.random.default_rng(9)
    corpus = rng.standard_normal((1500, 16)).astype(np.float32)
    ids = np.arange(1500, dtype=np.int64)[::-1].copy()
    q = rng.standard_normal((50, 16)).astype(np.float32)
    q[7, 3] = np.nan
    for ct in (0, 400):                                       # 400 rows per corpus tile < k: short per-tile lists
        gi, gd, gc = knn_join(q, ids, corpus, L2, 600, query_tile=16, corpus_tile_rows=ct)
        oi, od, oc = oracle.query_canonical(oracle.L2, corpus, ids, q, 600)
        assert (gc == oc).all() and (gi == oi).all()
        assert (onp.float_order_key(gd) == onp.float_order_key(od)).all()
    gi, gd, gc = knn_join(q, ids, corpus, L2, 0)
    assert gi.shape == (50, 0) and gc.tolist() == [0] * 50
    gi, gd, gc = knn_join(q, ids[:0], corpus[:0], L2, 5)
    assert (gi == -1).all() and np.isinf(gd).all() and gc.tolist() == [0] * 50
    same = np.ones((30_000, 8), np.float32)                   # every row ties: far more than the selector holds
    sid = rng.permutation(30_000).astype(np.int64)
    gi, gd, gc = knn_join(np.ones((3, 8), np.float32), sid, same, InnerProduct, 10, corpus_tile_rows=12_000)
    assert gi.tolist() == [list(range(10))] * 3 and gd.tolist() == [[-7.0] * 10] * 3 and gc.tolist() == [10] * 3
    with pytest.raises(_capi.AnnError) as e:
        knn_join(q, ids, corpus, L2, -1)
    assert e.value.code == _capi.ANN_ERR_NEGATIVE_K


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["thrift", "raw"])
@pytest.mark.parametrize("metric_name", ["L2", "Cosine", "InnerProduct"])
def test_serializable_index_round_trip(tmp_path, metric_name, fmt):
    from the_algorithm_b200.ann.brute_force import BruteForceIndex, SerializableBruteForceIndex
    from the_algorithm_b200.ann.common import FuturePool, Metric

    metric = Metric.from_string(metric_name)
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((5000, 37)).astype(np.float32)
    ids = rng.permutation(5000).astype(np.int64) - 2500
    q = rng.standard_normal((9, 37)).astype(np.float32)
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool())
    ix.append_batch(ids[:3000], rows[:3000])
    ix.append_batch(ids[3000:], rows[3000:])
    got_ids, got_rows = ix.read_rows(2990, 20)
    assert (got_ids == ids[2990:3010]).all() and (got_rows == rows[2990:3010]).all()
    SerializableBruteForceIndex.to_directory(ix, tmp_path / "idx", chunk_rows=1024, fmt=fmt)
    assert (tmp_path / "idx" / "BruteForceFileData").exists() and (tmp_path / "idx" / "_SUCCESS").exists()
    back = SerializableBruteForceIndex.from_directory(tmp_path / "idx", metric, FuturePool.immediate_pool(), chunk_rows=777)
    assert back.size() == 5000
    a, b = ix.batch_query_with_distance(q, 50), back.batch_query_with_distance(q, 50)
    assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()
    if fmt == "raw":     # the native raw layout records its metric; the reference's thrift stream does not (serialization.thrift:7-10)
        with pytest.raises(ValueError):
            SerializableBruteForceIndex.from_directory(tmp_path / "idx", Metric.from_thrift((metric.ordinal + 1) % 3),
                                                       FuturePool.immediate_pool())
    ix.close()
    back.close()
