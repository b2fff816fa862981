"""torchrun worker: 10M x 512 gallery row-sharded over WORLD_SIZE GPUs, 4096-query top-10, timed for 1 / 2 / 4 query chunks
(the exchange of chunk c runs under the scan of chunk c+1).  Rank 0 prints one line per setting.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/knn_scale_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from fire_b200 import _lib
from fire_b200.dist import ShardedGallery, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.init(local)
dev = torch.device("cuda", local)
N, D, Q, k = 10_000_000, 512, 4096, 10
lo, hi = shard_bounds(N, world, rank)
gal = ShardedGallery(D, hi - lo, rank, world, device=local)
gal.id_offset, gal.total, gal.bulk_total = lo, N, N
for c0 in range(lo, hi, 1_000_000):
    g = torch.Generator(device=dev); g.manual_seed(1000 + c0)
    gal.local.add(torch.randn(min(hi, c0 + 1_000_000) - c0, D, generator=g, device=dev))
gq = torch.Generator(device=dev); gq.manual_seed(4)
q = torch.randn(Q, D, generator=gq, device=dev)
ref = None
for chunks in (1, 2, 4, 1, 2):
    for _ in range(3):
        d, i = gal.search(q, k, chunks=chunks)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d, i = gal.search(q, k, chunks=chunks)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    same = True if ref is None else bool(torch.equal(i, ref))
    ref = i.clone() if ref is None else ref
    if rank == 0:
        ms = float(t.item())
        print(f"world {world} chunks {chunks}: {ms:.3f} ms/batch = {Q / ms * 1e3:.0f} QPS, per-GPU {2.0 * Q * N * D / ms / 1e9 / world:.0f} TFLOP/s, same ids {same}", flush=True)
dist.barrier()
dist.destroy_process_group()
