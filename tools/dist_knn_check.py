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
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_KNN_OK" if flag.item() == 1 else "DIST_KNN_MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
