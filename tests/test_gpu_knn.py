"""GPU parity: exact cosine top-k (fire_knn_* through the C ABI) vs the hnswlib-BFIndex oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(idx, ora, q, k, atol=5e-6):
    dist, ids = idx.search(q, k)
    ol, od = ora.knn_query(q, k, num_threads=8)
    ol = ol.astype(np.int64)
    assert dist.shape == od.shape and ids.dtype == np.int64 and dist.dtype == np.float32
    assert np.abs(dist - od).max() < atol                       # fp32 distances, different summation order
    assert np.all(np.diff(dist, axis=1) >= 0)                   # ascending
    bad = np.argwhere(ids != ol)
    for qi, j in bad:                                           # ids bit-exact, ties within 1e-5 excepted
        near = [od[qi, jj] for jj in (j - 1, j + 1) if 0 <= jj < k]
        assert any(abs(od[qi, j] - v) < 1e-5 for v in near), (qi, j, ids[qi], ol[qi], od[qi])
    return dist, ids, len(bad)


@pytest.mark.parametrize("N,D,Q,k", [(1000, 128, 5, 1), (300, 128, 3, 10), (5000, 512, 130, 10), (257, 512, 1, 1),
                                     (10000, 128, 32, 1), (20000, 512, 129, 10), (30000, 512, 40, 50), (70, 128, 9, 50),
                                     (4096, 256, 17, 10), (12, 128, 4, 10)])
def test_knn_matches_oracle(fire_lib, oracle_native, N, D, Q, k):
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(N + Q)
    g = rng.standard_normal((N, D), dtype=np.float32) * rng.uniform(0.1, 10, (N, 1)).astype(np.float32)
    q = rng.standard_normal((Q, D), dtype=np.float32)
    idx = KnnIndex(D, capacity=N + 10)
    idx.add(g[: N // 2]); idx.add(g[N // 2:])                  # two appends == one
    ora = oracle_native.BFIndexOracle(D); ora.add_items(g)
    assert idx.count == N
    _check(idx, ora, q, k)
    np.testing.assert_allclose(idx.rows(), ora.rows, atol=2e-7)  # stored rows are the hnswlib-normalised ones


def test_knn_duplicates_zero_vectors_and_tie_rule(fire_lib, oracle_native):
    """Ties: (distance asc, label asc) like BFIndex's max-heap of pairs; zero vector -> distance exactly 1."""
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(7)
    D = 128
    base = rng.standard_normal((50, D), dtype=np.float32)
    g = np.concatenate([base, base[:20], np.zeros((3, D), np.float32), base[:5] * 3.0])   # exact duplicates, zeros, scaled copies
    idx = KnnIndex(D, capacity=len(g)); idx.add(g)
    ora = oracle_native.BFIndexOracle(D); ora.add_items(g)
    q = np.concatenate([base[:10], np.zeros((1, D), np.float32)])
    dist, ids, _ = _check(idx, ora, q, 10)                      # scaled copies differ by 1 ulp after normalisation: 1e-5 tie window
    for i in range(10):                                         # bit-identical rows tie EXACTLY: lower id first, adjacent
        row = list(ids[i])
        assert row.index(i) + 1 == row.index(50 + i) and dist[i, row.index(i)] == dist[i, row.index(50 + i)]
        assert set(row[:3]) >= {i, 50 + i} and abs(dist[i, 0]) < 1e-6
    assert np.all(dist[-1] == 1.0)                              # zero query: every distance is exactly 1
    assert list(ids[-1]) == list(range(10))                     # ... and the all-way tie resolves by ascending id
    ol, _ = ora.knn_query(q[-1:], 10)
    assert list(ol[0]) == list(range(10))


def test_knn_forced_fallback_is_exact(fire_lib, oracle_native):
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(11)
    g = rng.standard_normal((20000, 512), dtype=np.float32)
    q = rng.standard_normal((300, 512), dtype=np.float32)
    idx = KnnIndex(512, capacity=20000); idx.add(g)
    idx.set_margin(0.5)                                         # every query fails the filter proof -> exact fp32 scan
    ora = oracle_native.BFIndexOracle(512); ora.add_items(g)
    _check(idx, ora, q, 10)
    total, fb = idx.stats()
    assert total == 300 and fb == 300                           # > EXACT_CAP exercises the overflow kernel too


def test_knn_near_duplicate_gallery_uses_proof(fire_lib, oracle_native):
    """A clustered gallery (many rows within 1e-3 cosine of each other) must flag queries and still be exact."""
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(13)
    D = 512
    centers = rng.standard_normal((20, D), dtype=np.float32)
    g = (centers[rng.integers(0, 20, 5000)] + 0.002 * rng.standard_normal((5000, D), dtype=np.float32)).astype(np.float32)
    q = centers + 0.002 * rng.standard_normal((20, D), dtype=np.float32)
    idx = KnnIndex(D, capacity=5000); idx.add(g)
    ora = oracle_native.BFIndexOracle(D); ora.add_items(g)
    _check(idx, ora, q.astype(np.float32), 10)
    assert idx.stats()[1] > 0


def test_logical_shards_merge_equals_unsharded(fire_lib, oracle_native):
    """Row-sharded search + fire_knn_merge == single-index search (the multi-GPU path on one device)."""
    import torch
    from fire_b200.engine import KnnIndex, knn_merge
    from fire_b200.dist import shard_bounds
    rng = np.random.default_rng(17)
    N, D, Q, k, G = 30011, 128, 77, 10, 4
    g = rng.standard_normal((N, D), dtype=np.float32)
    q = torch.from_numpy(rng.standard_normal((Q, D), dtype=np.float32)).cuda()
    full = KnnIndex(D, N); full.add(g)
    fd, fi = full.search(q, k)
    parts_d, parts_i = [], []
    for r in range(G):
        lo, hi = shard_bounds(N, G, r)
        sh = KnnIndex(D, hi - lo); sh.add(g[lo:hi])
        d, i = sh.search(q, k, id_offset=lo)
        parts_d.append(d); parts_i.append(i)
    md, mi = knn_merge(torch.stack(parts_d).contiguous(), torch.stack(parts_i).contiguous())
    torch.cuda.synchronize()
    assert torch.equal(mi, fi) and torch.equal(md, fd)


def test_knn_1m_properties(fire_lib, oracle_native):
    """BASELINE configs[2] size (1M x 512, Q=4096, k=10): size-independent properties + oracle on a query subset."""
    import torch
    from fire_b200.engine import KnnIndex
    N, D, Q, k = 1_000_000, 512, 4096, 10
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    g = torch.randn(N, D, generator=gen, device="cuda")
    idx = KnnIndex(D, N); idx.add(g)
    rows = torch.randint(0, N, (Q,), generator=gen, device="cuda")
    q = g[rows].contiguous()                                    # self-queries: the row itself must come back first
    dist, ids = idx.search(q, k)
    torch.cuda.synchronize()
    assert torch.all(ids[:, 0] == rows) and float(dist[:, 0].abs().max()) < 1e-6
    assert torch.all(dist[:, 1:] >= dist[:, :-1]) and int(ids.min()) >= 0 and int(ids.max()) < N
    # oracle on 24 fresh queries against the first 200k rows (CPU brute force stays in seconds)
    sub = KnnIndex(D, 200_000); sub.add(g[:200_000].contiguous())
    q2 = torch.randn(24, D, generator=gen, device="cuda")
    d2, i2 = sub.search(q2, k)
    ora = oracle_native.BFIndexOracle(D); ora.add_items(g[:200_000].cpu().numpy())
    ol, od = ora.knn_query(q2.cpu().numpy(), k, num_threads=8)
    assert np.array_equal(i2.cpu().numpy(), ol.astype(np.int64)) and np.abs(d2.cpu().numpy() - od).max() < 5e-6
    assert idx.stats()[1] / idx.stats()[0] < 0.02               # the fp16 filter proof almost never fails on random data


def test_knn_errors(fire_lib):
    from fire_b200.engine import KnnIndex
    from fire_b200._lib import FireError
    idx = KnnIndex(128, 10)
    idx.add(np.ones((3, 128), np.float32))
    with pytest.raises(FireError):
        idx.search(np.ones((1, 128), np.float32), 5)           # k > count (hnswlib raises too)
    with pytest.raises(FireError):
        idx.add(np.ones((8, 128), np.float32))                 # capacity
    with pytest.raises(FireError):
        KnnIndex(100, 10)                                      # D must be a multiple of 64
