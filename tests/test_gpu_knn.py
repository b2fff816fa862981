"""GPU parity: exact cosine top-k (fire_knn_* through the C ABI) vs the hnswlib-BFIndex oracle."""
import os

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
    # the full config against the oracle: a 256-query subset of a fresh 4096-query batch vs ALL 1M rows
    q2 = torch.randn(Q, D, generator=gen, device="cuda")
    d2, i2 = idx.search(q2, k)
    sel = torch.arange(0, Q, Q // 256, device="cuda")[:256]
    ora = oracle_native.BFIndexOracle(D)
    ora.rows = idx.rows()                                       # the stored rows ARE the hnswlib-normalised ones (test_knn_matches_oracle)
    ora.labels = np.arange(N, dtype=np.uint64)
    ol, od = ora.knn_query(q2[sel].cpu().numpy(), k, num_threads=os.cpu_count() or 8)
    got_i, got_d = i2[sel].cpu().numpy(), d2[sel].cpu().numpy()
    assert np.abs(got_d - od).max() < 5e-6
    for qi, j in np.argwhere(got_i != ol.astype(np.int64)):     # ids bit-exact, ties within 1e-5 excepted
        near = [od[qi, jj] for jj in (j - 1, j + 1) if 0 <= jj < k]
        assert any(abs(od[qi, j] - v) < 1e-5 for v in near), (qi, j, got_i[qi], ol[qi], od[qi])
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


def test_knn_add_search_add_search_host_path(fire_lib, oracle_native):
    """add(small) -> search -> add(larger, regrows the staging buffer) -> search through the numpy entry points: the
    regrow must not touch the search buffers (round-1 advisor finding: a stray cudaFree in fire_knn_add_host)."""
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(23)
    D = 128
    g = rng.standard_normal((700, D), dtype=np.float32)
    q = rng.standard_normal((9, D), dtype=np.float32)
    idx = KnnIndex(D, 1000)
    ora = oracle_native.BFIndexOracle(D)
    for lo, hi in ((0, 10), (10, 30), (30, 700)):               # every add is larger than the one before
        idx.add(g[lo:hi]); ora.add_items(g[lo:hi])
        _check(idx, ora, q, min(10, hi))
        _check(idx, ora, q[:3], 1)


def test_knn_packed_records_and_short_shards(fire_lib, oracle_native):
    """fire_knn_search_packed + fire_knn_merge_packed (ONE buffer for the multi-GPU exchange) == the two-array path; shards
    that hold fewer than k rows (or none) contribute padding and the merged result is still the unsharded oracle's."""
    import torch
    from fire_b200.engine import KnnIndex, knn_merge_packed
    rng = np.random.default_rng(29)
    D, Q, k, G = 128, 33, 10, 4
    for N, layout in ((30011, "contiguous"), (30011, "interleaved"), (23, "interleaved"), (3, "interleaved"), (11, "contiguous")):
        g = rng.standard_normal((N, D), dtype=np.float32)
        q = torch.from_numpy(rng.standard_normal((Q, D), dtype=np.float32)).cuda()
        kk = min(k, N)
        recs = []
        for r in range(G):
            if layout == "contiguous":
                lo, hi = (r * N) // G, ((r + 1) * N) // G
                rows, off, stride = g[lo:hi], lo, 1
            else:
                rows, off, stride = g[r::G], r, G
            sh = KnnIndex(D, max(1, len(rows)))
            if len(rows):
                sh.add(np.ascontiguousarray(rows))
            recs.append(sh.search_packed(q, kk, off, stride))
        md, mi = knn_merge_packed(torch.stack(recs).contiguous())
        torch.cuda.synchronize()
        ora = oracle_native.BFIndexOracle(D); ora.add_items(g)
        ol, od = ora.knn_query(q.cpu().numpy(), kk)
        assert np.array_equal(mi.cpu().numpy(), ol.astype(np.int64)), (N, layout)
        assert np.abs(md.cpu().numpy() - od).max() < 5e-6


def test_knn_search_rows_is_bulk_find_similar(fire_lib, oracle_native):
    """fire_knn_search_rows: the stored rows as queries == knn_query(original vector) for every row (hnsw_manager.py:227-244)."""
    import torch
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(31)
    N, D, k = 3000, 128, 50
    centers = rng.standard_normal((300, D), dtype=np.float32)
    g = (centers[rng.integers(0, 300, N)] + 0.3 * rng.standard_normal((N, D), dtype=np.float32)).astype(np.float32)
    idx = KnnIndex(D, N); idx.add(g)
    ora = oracle_native.BFIndexOracle(D); ora.add_items(g)
    ol, od = ora.knn_query(g, k, num_threads=8)
    for first, n in ((0, 1000), (1000, 2000)):
        d, i = idx.search_rows(first, n, k)
        torch.cuda.synchronize()
        d, i = d.cpu().numpy(), i.cpu().numpy()
        assert np.abs(d - od[first:first + n]).max() < 5e-6
        for qi, j in np.argwhere(i != ol[first:first + n].astype(np.int64)):
            near = [od[first + qi, jj] for jj in (j - 1, j + 1) if 0 <= jj < k]
            assert any(abs(od[first + qi, j] - v) < 1e-5 for v in near)
        assert np.array_equal(i[:, 0], np.arange(first, first + n))          # every row finds itself first


def test_knn_bounded_fallback_paths_are_exact(fire_lib, oracle_native):
    """The three exits of a flagged query - candidates only, ONE gallery split re-scanned, whole shard re-scanned - all
    return the oracle's ids; the margin knob moves queries between them."""
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(37)
    N, D, Q, k = 60000, 512, 200, 10
    g = rng.standard_normal((N, D), dtype=np.float32)
    q = rng.standard_normal((Q, D), dtype=np.float32)
    ora = oracle_native.BFIndexOracle(D); ora.add_items(g)
    seen = np.zeros(3, dtype=np.int64)
    for eps in (0.004, 0.03, 0.052, 0.056, 0.09):      # vs the ~0.006 gap to the merged KP-th score and ~0.055 to one split's own
        idx = KnnIndex(D, N); idx.add(g); idx.set_margin(eps)
        _check(idx, ora, q, k)
        total, flagged, one_split, whole = idx.stats_ex()
        assert total == Q and flagged >= one_split + whole
        seen += np.array([flagged - one_split - whole, one_split, whole])
        idx.close()
    assert np.all(seen > 0), seen                               # every exit was exercised at least once


def test_knn_nan_query_writes_every_slot(fire_lib):
    """A NaN query cannot leave output slots unwritten (round-1 advisor finding): padding is (FLT_MAX, -1)."""
    import torch
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(41)
    idx = KnnIndex(128, 500); idx.add(rng.standard_normal((500, 128), dtype=np.float32))
    q = torch.from_numpy(rng.standard_normal((4, 128), dtype=np.float32)).cuda()
    q[1, 5] = float("nan")
    d = torch.full((4, 10), 123.0, device="cuda"); i = torch.full((4, 10), 123456789, dtype=torch.int64, device="cuda")
    idx.search(q, 10, out_dist=d, out_ids=i)
    torch.cuda.synchronize()
    assert int((i == 123456789).sum()) == 0 and int((d == 123.0).sum()) == 0
    assert bool(((i[1] == -1) | ((i[1] >= 0) & (i[1] < 500))).all())
    assert bool((i[[0, 2, 3]] >= 0).all())


def test_knn_cta_pair_scan_is_identical(fire_lib, oracle_native, monkeypatch):
    """The cta_group::2 instantiation of knn_scan_kernel (FIRE_B200_KNN_PAIR=1; off by default: no gain measured,
    profiles/r02_knn_experiments.txt) returns exactly what the single-CTA kernel returns."""
    import torch
    from fire_b200.engine import KnnIndex
    rng = np.random.default_rng(43)
    g = rng.standard_normal((50000, 512), dtype=np.float32)
    idx = KnnIndex(512, 50000); idx.add(g)
    for Q, k in ((256, 10), (130, 1), (1024, 10), (200, 50)):
        q = torch.from_numpy(rng.standard_normal((Q, 512), dtype=np.float32)).cuda()
        monkeypatch.setenv("FIRE_B200_KNN_PAIR", "0")
        d0, i0 = idx.search(q, k)
        monkeypatch.setenv("FIRE_B200_KNN_PAIR", "1")
        d1, i1 = idx.search(q, k)
        torch.cuda.synchronize()
        assert torch.equal(i0, i1) and torch.equal(d0, d1), (Q, k)
