"""The HNSW restatement used for the recall report (oracle/hnsw_oracle.c): sanity properties only - it is PARITY
UNPINNED against hnswlib (the wheel is not installable here) and never gates the product."""
import numpy as np
import pytest


def _recall(a, b):
    return float(np.mean([len(set(a[i].tolist()) & set(b[i].tolist())) / a.shape[1] for i in range(len(a))]))


def test_hnsw_oracle_recall_and_ordering(oracle_native):
    rng = np.random.default_rng(3)
    g = rng.standard_normal((3000, 64)).astype(np.float32)
    q = rng.standard_normal((100, 64)).astype(np.float32)
    h = oracle_native.HnswOracle(64, 3000)
    h.add_items(g)
    assert h.get_current_count() == 3000
    bf = oracle_native.BFIndexOracle(64)
    bf.add_items(g)
    el, ed = bf.knn_query(q, 10)
    h.set_ef(200)
    l, d = h.knn_query(q, 10)
    assert l.dtype == np.uint64 and d.dtype == np.float32 and l.shape == (100, 10)
    assert (np.diff(d, axis=1) >= 0).all()                       # ascending distances like hnswlib
    assert _recall(l.astype(np.int64), el.astype(np.int64)) >= 0.9
    h.set_ef(3000)                                               # search width = whole index: the graph search is exhaustive
    l2, d2 = h.knn_query(q, 10)
    assert _recall(l2.astype(np.int64), el.astype(np.int64)) >= 0.995
    hit = l2.astype(np.int64) == el.astype(np.int64)
    assert np.abs(d2[hit] - ed[hit]).max() < 1e-5                # same distance definition: 1 - <q^, g^>
    h.set_ef(10)                                                 # the reference's accidental setting after load_index
    l3, _ = h.knn_query(q, 10)
    assert _recall(l3.astype(np.int64), el.astype(np.int64)) < _recall(l.astype(np.int64), el.astype(np.int64))


def test_hnsw_oracle_k_larger_than_count_raises(oracle_native):
    h = oracle_native.HnswOracle(8, 100)
    h.add_items(np.eye(8, dtype=np.float32)[:3])
    with pytest.raises(RuntimeError):
        h.knn_query(np.ones((1, 8), dtype=np.float32), 5)
    l, d = h.knn_query(np.eye(8, dtype=np.float32)[1:2], 1)
    assert int(l[0, 0]) == 1 and abs(float(d[0, 0])) < 1e-6
