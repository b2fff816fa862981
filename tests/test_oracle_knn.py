"""Known-answer tests for the hnswlib BFIndex restatement (oracle/fire_oracle.c).  hnswlib itself cannot be installed
here, so these pin the semantics stated in SURVEY App. B: cosine normalisation with 1/(||x||+1e-30), distance
1 - <q,g>, result = k lexicographically smallest (distance, label), ascending."""
import numpy as np
import pytest


def test_identity_gallery(oracle_native):
    ora = oracle_native.BFIndexOracle(4)
    ora.add_items(np.eye(4, dtype=np.float32) * np.array([[1], [2], [3], [4]], np.float32))
    labels, dist = ora.knn_query(np.array([[0, 5, 0, 0]], np.float32), 4)
    assert labels.dtype == np.uint64 and dist.dtype == np.float32
    assert list(labels[0]) == [1, 0, 2, 3]                           # exact match first, then ties (dist 1) by label
    assert dist[0, 0] == 0.0 and np.all(dist[0, 1:] == 1.0)


def test_duplicate_rows_tie_by_label_and_custom_ids(oracle_native):
    ora = oracle_native.BFIndexOracle(16)
    v = np.arange(1, 17, dtype=np.float32)
    ora.add_items(np.stack([v, v * 2, -v, v]), ids=[40, 10, 30, 20])
    labels, dist = ora.knn_query(v, 3)
    assert list(labels[0]) == [10, 20, 40]                           # three identical directions: ascending label
    assert np.all(np.abs(dist[0]) < 1e-6)
    labels, dist = ora.knn_query(v, 4)
    assert labels[0, 3] == 30 and abs(dist[0, 3] - 2.0) < 1e-6


def test_zero_vector_and_k_equals_count(oracle_native):
    ora = oracle_native.BFIndexOracle(16)
    rng = np.random.default_rng(0)
    g = rng.standard_normal((5, 16), dtype=np.float32)
    g[2] = 0
    ora.add_items(g)
    assert not np.isnan(ora.rows).any() and not ora.rows[2].any()    # zero stays zero (no NaN)
    labels, dist = ora.knn_query(np.zeros(16, np.float32), 5)
    assert list(labels[0]) == [0, 1, 2, 3, 4] and np.all(dist == 1.0)
    with pytest.raises(RuntimeError):
        ora.knn_query(g[0], 6)                                       # hnswlib: "Cannot return the results in a contiguous 2D array"


@pytest.mark.parametrize("D,k", [(128, 1), (128, 10), (512, 10), (512, 50), (24, 3)])
def test_against_float64_bruteforce(oracle_native, D, k):
    rng = np.random.default_rng(D + k)
    g = rng.standard_normal((3000, D), dtype=np.float32) * 3
    q = rng.standard_normal((40, D), dtype=np.float32)
    ora = oracle_native.BFIndexOracle(D)
    ora.add_items(g)
    labels, dist = ora.knn_query(q, k, num_threads=4)
    gn = g.astype(np.float64) / np.linalg.norm(g.astype(np.float64), axis=1, keepdims=True)
    qn = q.astype(np.float64) / np.linalg.norm(q.astype(np.float64), axis=1, keepdims=True)
    d64 = 1.0 - qn @ gn.T
    order = np.argsort(d64, axis=1, kind="stable")[:, :k]
    assert np.abs(np.take_along_axis(d64, order, 1) - dist).max() < 2e-6
    mism = labels.astype(np.int64) != order
    assert mism.mean() < 0.01                                        # only fp32-vs-fp64 near-ties may differ
    assert np.all(np.diff(dist, axis=1) >= 0)
    one, _ = ora.knn_query(q, k, num_threads=1)
    assert np.array_equal(one, labels)                               # threading over queries does not change results


@pytest.mark.parametrize("D,k", [(128, 10), (512, 50)])
def test_against_scikit_learn_bruteforce(oracle_native, D, k):
    """An executor the builder did not write: scikit-learn's exact cosine search (`NearestNeighbors(algorithm="brute",
    metric="cosine")`, distance = 1 - cos like hnswlib's cosine space).  It pins the semantics - which rows, in which order, at
    which distance - not hnswlib's summation order (that wheel is not installable here)."""
    sk = pytest.importorskip("sklearn.neighbors")
    rng = np.random.default_rng(1000 + D)
    g = rng.standard_normal((4000, D)).astype(np.float32) * 2
    q = rng.standard_normal((64, D)).astype(np.float32)
    ora = oracle_native.BFIndexOracle(D)
    ora.add_items(g)
    labels, dist = ora.knn_query(q, k, num_threads=4)
    nn = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(g.astype(np.float64))
    d_sk, i_sk = nn.kneighbors(q.astype(np.float64))
    assert np.abs(d_sk - dist).max() < 2e-6
    assert (labels.astype(np.int64) != i_sk).mean() < 0.01           # only fp32-vs-fp64 near-ties may differ
    gap_ok = np.abs(np.take_along_axis(d_sk, np.argsort(d_sk, 1), 1) - d_sk).max() == 0   # scikit-learn returns ascending distances too
    assert gap_ok


def test_normalize_formula(oracle_native):
    x = np.array([[3.0, 4.0] + [0.0] * 14], np.float32)
    y = oracle_native.normalize(x)
    assert abs(y[0, 0] - 0.6) < 1e-7 and abs(y[0, 1] - 0.8) < 1e-7


def test_blocked_scan_is_bit_identical(oracle_native):
    """The query-blocked loop order used by the BASELINE-sized parity checks changes nothing but the DRAM traffic."""
    rng = np.random.default_rng(77)
    g = rng.standard_normal((5000, 128)).astype(np.float32)
    g[100] = g[7]; g[4000] = g[7]                                # exact ties across the scan
    q = rng.standard_normal((37, 128)).astype(np.float32)
    q[3] = g[7]
    ora = oracle_native.BFIndexOracle(128); ora.add_items(g)
    for k in (1, 10, 50):
        for threads in (1, 3, 8):
            a = ora.knn_query(q, k, num_threads=threads)
            b = ora.knn_query(q, k, num_threads=threads, blocked=True)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
