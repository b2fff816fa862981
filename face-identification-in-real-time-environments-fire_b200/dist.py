"""Multi-GPU layer: one process per GPU (torchrun), torch.distributed for the plumbing.

SURVEY §8(e): the identification path shards two ways and has exactly one exchange step.

  * crops / frames are independent  -> data parallel, no collective (`shard_bounds` picks each rank's slice);
  * the gallery shards row-wise     -> every rank searches its contiguous slice for ALL queries (exact local
    top-k, ids offset by the slice start), the per-rank (distance, id) lists are exchanged with ONE
    all_gather (NCCL over NVLink/NVSwitch; 12 bytes x Q x k per rank - latency bound), and every rank
    merges the G lists with the same (distance asc, id asc) rule, so all ranks hold the identical result.

`ShardedGallery` takes its local index and merge function as parameters: on the GPU they are
fire_b200.engine.KnnIndex / knn_merge (the default); the gloo CPU tests inject test doubles so the
partitioning, offsets and collective plumbing are covered without a GPU.  There is no CPU fallback
in the product: the defaults raise without a B200.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n rows over `world` ranks: rank r owns [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedGallery:
    def __init__(self, dim: int, capacity_per_rank: int, rank: int = 0, world: int = 1, device: int = 0,
                 local_index=None, merge_fn: Optional[Callable] = None, group=None):
        self.dim, self.rank, self.world, self.group = dim, rank, world, group
        if local_index is None:
            from .engine import KnnIndex
            local_index = KnnIndex(dim, capacity=capacity_per_rank, device=device)
        if merge_fn is None:
            from .engine import knn_merge
            merge_fn = knn_merge
        self.local = local_index
        self.merge_fn = merge_fn
        self.id_offset = 0
        self.total = 0

    def add_global(self, n_total: int, rows_for_range: Callable[[int, int], object]):
        """Enrol rows [0, n_total): this rank materialises and stores only its own slice.
        rows_for_range(lo, hi) returns the rows of that slice (numpy on the host path, cuda tensor on the device path)."""
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        self.id_offset = lo
        self.total = n_total
        if hi > lo:
            self.local.add(rows_for_range(lo, hi))

    def search(self, queries, k: int):
        """queries: replicated on every rank.  Returns (dist [Q,k], ids [Q,k]) - identical on all ranks."""
        import torch
        import torch.distributed as dist
        d, i = self.local.search(queries, k, id_offset=self.id_offset)
        if self.world == 1:
            return d, i
        d = d if torch.is_tensor(d) else torch.from_numpy(np.ascontiguousarray(d))
        i = i if torch.is_tensor(i) else torch.from_numpy(np.ascontiguousarray(i))
        Q, k = d.shape
        # rank-major concatenation == [G][Q][k] in memory (the layout fire_knn_merge reads); gloo accepts only this form
        gd = torch.empty((self.world * Q, k), dtype=d.dtype, device=d.device)
        gi = torch.empty((self.world * Q, k), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gd, d.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi, i.contiguous(), group=self.group)
        return self.merge_fn(gd.view(self.world, Q, k), gi.view(self.world, Q, k))


def gather_embeddings(local_emb, world: int, group=None):
    """DP encode followed by a gallery search needs every rank's embeddings: one all_gather of [B/G, D]."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_emb
    out = torch.empty((world * local_emb.shape[0],) + tuple(local_emb.shape[1:]), dtype=local_emb.dtype,
                      device=local_emb.device)
    dist.all_gather_into_tensor(out, local_emb.contiguous(), group=group)
    return out
