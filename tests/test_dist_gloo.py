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
    gal = ShardedGallery(D, N, rank, world, local_index=fakes.FakeKnnIndex(D, N), merge_fn=fakes.numpy_merge)
    gal.add_global(N, lambda lo, hi: g[lo:hi])
    assert gal.id_offset == shard_bounds(N, world, rank)[0] and gal.local.count == shard_bounds(N, world, rank)[1] - gal.id_offset
    d, i = gal.search(torch.from_numpy(q), k)
    emb = gather_embeddings(torch.full((3, 4), float(rank)), world)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), d=d.numpy(), i=i.numpy(), emb=emb.numpy())
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
