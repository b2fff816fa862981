"""torchrun worker: row-sharded gallery over WORLD_SIZE GPUs (NCCL all_gather + fire_knn_merge) must equal the
single-GPU search of the whole gallery.  Prints DIST_KNN_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from fire_b200 import _lib
from fire_b200.dist import ShardedGallery
from fire_b200.engine import KnnIndex

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.init(local)
N, D, Q, k = 300_007, 512, 777, 10
g = torch.Generator(device="cuda"); g.manual_seed(99)
rows = torch.randn(N, D, generator=g, device="cuda")          # same seed on every rank -> same gallery
rows[200_000] = rows[17]                                        # cross-shard exact tie
q = torch.randn(Q, D, generator=g, device="cuda"); q[0] = rows[17]
gal = ShardedGallery(D, N, rank, world, device=local)
gal.add_global(N, lambda lo, hi: rows[lo:hi].contiguous())
d, i = gal.search(q, k)
full = KnnIndex(D, N, device=local); full.add(rows)
fd, fi = full.search(q, k)
torch.cuda.synchronize()
ok = torch.equal(i, fi) and torch.equal(d, fd) and i[0, 0].item() == 17 and i[0, 1].item() == 200_000
# the chunk-pipelined exchange (all_gather of chunk c under the scan of chunk c+1) gives the same answer
for chunks in (1, 3):
    d3, i3 = gal.search(q, k, chunks=chunks)
    torch.cuda.synchronize()
    ok = ok and torch.equal(i3, fi) and torch.equal(d3, fd)
# a gallery that GROWS across the ranks one row at a time (hnsw_manager.py:135-143 sharded): interleaved layout, from empty;
# early on some shard holds fewer than k rows (or none) and pads its lists
grow = ShardedGallery(D, 64, rank, world, device=local, layout="interleaved")
single = KnnIndex(D, 64, device=local)
for n in range(1, 40):
    row = rows[n - 1:n].clone()
    assert grow.add_embedding(grow.broadcast_row(row)) == n - 1
    single.add(row)
    if n in (1, 2, 3, 11, 39):
        kk = min(k, n)
        gd, gi = grow.search(q[:9].contiguous(), kk)
        sd, si = single.search(q[:9].contiguous(), kk)
        torch.cuda.synchronize()
        ok = ok and torch.equal(gi, si) and torch.equal(gd, sd)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_KNN_OK" if flag.item() == 1 else "DIST_KNN_MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
