"""Multi-GPU host logic on CPU: world_size-2 `gloo` processes run fire_b200.dist.ShardedGallery with oracle-backed
test doubles for the local index and the merge kernel; the sharded result must equal the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from fire_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import fakes
    from fire_b200.dist import ShardedGallery, gather_embeddings, shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                       # same data on every rank; each enrols only its slice
    N, D, Q, k = 4001, 32, 37, 10
    g = rng.standard_normal((N, D), dtype=np.float32)
    g[2000] = g[5]                                         # a cross-shard exact tie: must resolve to the lower id
    q = rng.standard_normal((Q, D), dtype=np.float32)
    q[0] = g[5]
    gal = ShardedGallery(D, N, rank, world, local_index=fakes.FakeKnnIndex(D, N), merge_fn=fakes.numpy_merge_packed)
    gal.add_global(N, lambda lo, hi: g[lo:hi])
    assert gal.id_offset == shard_bounds(N, world, rank)[0] and gal.local.count == shard_bounds(N, world, rank)[1] - gal.id_offset
    d, i = gal.search(torch.from_numpy(q), k)
    d3, i3 = gal.search(torch.from_numpy(q), k, chunks=3)             # chunk-pipelined exchange: same answer
    assert torch.equal(d, d3) and torch.equal(i, i3)
    emb = gather_embeddings(torch.full((3, 4), float(rank)), world)
    out = dict(d=d.numpy(), i=i.numpy(), emb=emb.numpy())

    # interleaved layout grown one row at a time (hnsw_manager.py:135-143 across shards), starting EMPTY: while the
    # gallery holds fewer rows than world * k (or fewer than `world`) some shard is short or empty and pads its lists
    gi = ShardedGallery(D, 64, rank, world, local_index=fakes.FakeKnnIndex(D, 64), merge_fn=fakes.numpy_merge_packed, layout="interleaved")
    try:
        gi.search(torch.from_numpy(q[:2]), 1)
        raise AssertionError("k > count must raise like hnswlib")
    except RuntimeError as e:
        assert "contiguous 2D array" in str(e)
    grown = []
    for n in range(1, 26):
        new_id = gi.add_embedding(gi.broadcast_row(torch.from_numpy(g[n - 1].copy())).numpy())
        assert new_id == n - 1 and gi.total == n
        if n in (1, 2, 3, 11, 25):
            kk = min(10, n)
            dd, ii = gi.search(torch.from_numpy(q[:5]), kk)
            grown.append((n, dd.numpy(), ii.numpy()))
    assert gi.local.count == (25 - rank + world - 1) // world
    for n, dd, ii in grown:
        out[f"gd{n}"], out[f"gi{n}"] = dd, ii
    # contiguous layout: appended rows extend the last shard
    gc = ShardedGallery(D, 40, rank, world, local_index=fakes.FakeKnnIndex(D, 40), merge_fn=fakes.numpy_merge_packed)
    gc.add_global(21, lambda lo, hi: g[lo:hi])
    for n in range(21, 30):
        assert gc.add_embedding(g[n]) == n
    dd, ii = gc.search(torch.from_numpy(q[:5]), 10)
    out["cd"], out["ci"] = dd.numpy(), ii.numpy()
    # a full owner shard raises on EVERY rank (nobody is left inside a collective)
    gf = ShardedGallery(D, 1, rank, world, local_index=fakes.FakeKnnIndex(D, 1), merge_fn=fakes.numpy_merge_packed, layout="interleaved")
    gf.add_embedding(g[0]); gf.add_embedding(g[1])
    try:
        gf.add_embedding(g[2])
        raise AssertionError("full shard must raise")
    except RuntimeError as e:
        assert "full" in str(e)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_gallery_equals_unsharded_oracle(tmp_path, oracle_native):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    N, D, Q, k = 4001, 32, 37, 10
    g = rng.standard_normal((N, D), dtype=np.float32)
    g[2000] = g[5]
    q = rng.standard_normal((Q, D), dtype=np.float32)
    q[0] = g[5]
    ora = oracle_native.BFIndexOracle(D)
    ora.add_items(g)
    ol, od = ora.knn_query(q, k)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for o in outs:                                         # identical on every rank, equal to the unsharded oracle
        assert np.array_equal(o["i"], ol.astype(np.int64)) and np.array_equal(o["d"], od)
        assert list(o["i"][0][:2]) == [5, 2000]
        assert np.array_equal(o["emb"][:, 0], np.array([0, 0, 0, 1, 1, 1], np.float32))
        for n in (1, 2, 3, 11, 25):                        # the growing interleaved gallery == an unsharded index of the same rows
            o2 = oracle_native.BFIndexOracle(D)
            o2.add_items(g[:n])
            l2, d2 = o2.knn_query(q[:5], min(10, n))
            assert np.array_equal(o[f"gi{n}"], l2.astype(np.int64)) and np.array_equal(o[f"gd{n}"], d2), n
        o3 = oracle_native.BFIndexOracle(D)
        o3.add_items(g[:30])
        l3, d3 = o3.knn_query(q[:5], 10)
        assert np.array_equal(o["ci"], l3.astype(np.int64)) and np.array_equal(o["cd"], d3)
