"""Multi-GPU layer: one process per GPU (torchrun), torch.distributed for the plumbing.

SURVEY §8(e): the identification path shards two ways and has exactly one exchange step.

  * crops / frames are independent  -> data parallel, no collective (`shard_bounds` picks each rank's slice);
  * the gallery shards row-wise     -> every rank searches its own rows for ALL queries (exact local top-k with
    global ids), the per-rank lists are exchanged with ONE all_gather of packed 12-byte (distance, id) records
    (NCCL over NVLink/NVSwitch; 12 x Q x k bytes per rank - latency bound), and every rank merges the G lists with
    the same (distance asc, id asc) rule, so all ranks hold the identical result.  The query batch is cut into
    chunks so that the all-gather and merge of chunk i run under the scan of chunk i+1: only the last chunk's
    exchange is exposed;
  * enrolment (hnsw_manager.py:135-143 `add_embedding`) appends one row on the rank that owns the next id; labels
    and db ids stay host-side with the caller, like the reference's lists.

Two row layouts:
  "contiguous"  (BASELINE configs[3]): rank r holds rows [lo_r, hi_r) of a bulk-loaded gallery, id = lo_r + local row;
                appended rows go to the last rank.
  "interleaved": id g lives on rank g % G at local row g // G; bulk load and one-by-one enrolment both keep the
                shards balanced, so this is the layout of a gallery that grows while it is served.

`ShardedGallery` takes its local index and merge function as parameters: on the GPU they are
fire_b200.engine.KnnIndex / knn_merge_packed (the default); the gloo CPU tests inject test doubles so the
partitioning, id mapping, padding and collective plumbing are covered without a GPU.  There is no CPU fallback in
the product: the defaults raise without a B200.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import numpy as np

K_TOO_LARGE = "Cannot return the results in a contiguous 2D array. Probably ef or M is too small"   # hnswlib's message


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n rows over `world` ranks: rank r owns [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def interleaved_count(n: int, world: int, rank: int) -> int:
    """Rows of ids 0..n-1 that live on `rank` under the interleaved layout (id % world == rank)."""
    return (n - rank + world - 1) // world if n > rank else 0


class ShardedGallery:
    def __init__(self, dim: int, capacity_per_rank: int, rank: int = 0, world: int = 1, device: int = 0,
                 local_index=None, merge_fn: Optional[Callable] = None, group=None, layout: str = "contiguous",
                 chunks: Optional[int] = None):
        assert layout in ("contiguous", "interleaved")
        self.dim, self.rank, self.world, self.group, self.layout = dim, rank, world, group, layout
        if local_index is None:
            from .engine import KnnIndex
            local_index = KnnIndex(dim, capacity=capacity_per_rank, device=device)
        if merge_fn is None:
            from .engine import knn_merge_packed
            merge_fn = knn_merge_packed
        self.local = local_index
        self.merge_fn = merge_fn
        self.capacity_per_rank = capacity_per_rank
        self.id_offset = rank if layout == "interleaved" else 0
        self.id_stride = world if layout == "interleaved" else 1
        self.total = 0
        self.bulk_total = 0            # rows loaded by add_global (they fix the contiguous layout's shard boundaries)
        env = os.environ.get("FIRE_B200_KNN_CHUNKS")
        self.chunks = chunks if chunks is not None else (int(env) if env else None)

    # ---- enrolment -------------------------------------------------------------------------------------------------
    def add_global(self, n_total: int, rows_for_range: Callable[[int, int], object]):
        """Bulk-enrol ids [0, n_total) into an empty gallery: this rank materialises and stores only its own rows.
        rows_for_range(lo, hi) returns rows lo..hi-1 (numpy on the host path, cuda tensor on the device path); the
        interleaved layout asks for the whole range in bounded pieces and keeps every `world`-th row."""
        assert self.total == 0, "add_global loads an empty gallery; use add_embedding to append"
        if self.layout == "contiguous":
            lo, hi = shard_bounds(n_total, self.world, self.rank)
            self.id_offset = lo
            if hi > lo:
                self.local.add(rows_for_range(lo, hi))
        else:
            step = max(self.world, (1 << 18) // self.world * self.world)          # piece boundaries are multiples of world
            for lo in range(0, n_total, step):
                hi = min(n_total, lo + step)
                mine = rows_for_range(lo, hi)[self.rank::self.world]
                if len(mine):
                    self.local.add(mine if isinstance(mine, np.ndarray) else mine.contiguous())
        self.total = self.bulk_total = n_total

    def owner_of_next(self) -> int:
        return self.total % self.world if self.layout == "interleaved" else self.world - 1

    def add_embedding(self, row) -> int:
        """Append ONE row (the same `row` on every rank, e.g. after `broadcast_row`): the owner of the next id stores it
        (hnswlib-normalised by the index), every rank advances the count.  Returns the new row's global id.  Raises the
        same error on every rank when the owner's shard is full, so no rank is left waiting in a collective."""
        owner = self.owner_of_next()
        if self.layout == "interleaved":
            owner_count = interleaved_count(self.total, self.world, owner)
        else:                          # bulk rows are split by shard_bounds; every appended row sits behind the last rank's slice
            owner_count = self.total - shard_bounds(self.bulk_total, self.world, owner)[0]
        if owner_count >= self.capacity_per_rank:
            raise RuntimeError(f"sharded gallery: shard {owner} is full ({owner_count} rows)")
        new_id = self.total
        if self.rank == owner:
            self.local.add(row.reshape(1, self.dim))
        self.total += 1
        return new_id

    def broadcast_row(self, row, src: int = 0):
        """Convenience for enrolment: rank `src` holds the new embedding, everyone gets it (2 KB)."""
        import torch.distributed as dist
        if self.world > 1:
            dist.broadcast(row, src=src, group=self.group)
        return row

    # ---- search ----------------------------------------------------------------------------------------------------
    def search(self, queries, k: int, chunks: Optional[int] = None):
        """queries: replicated on every rank.  Returns (dist [Q,k], ids [Q,k]) - identical on all ranks.
        hnswlib's `k > count` rule applies to the GLOBAL row count; a shard with fewer than k rows (a small or empty
        gallery, N < world) contributes padded lists instead of failing."""
        import torch
        import torch.distributed as dist
        if k > self.total:
            raise RuntimeError(K_TOO_LARGE)
        if self.world == 1 and self.layout == "contiguous":
            return self.local.search(queries, k, id_offset=self.id_offset)
        if not torch.is_tensor(queries):
            queries = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))
        dev = getattr(self.local, "device", None)                  # the real KnnIndex lives on a GPU; test doubles do not
        if dev is not None and queries.device != dev:
            queries = queries.to(dev)
        queries = queries.reshape(-1, self.dim)
        Q = queries.shape[0]
        n_chunks = chunks if chunks is not None else self.chunks
        if n_chunks is None:
            n_chunks = 2 if (self.world > 1 and Q >= 2048) else 1
        n_chunks = max(1, min(n_chunks, Q))
        bounds = [shard_bounds(Q, n_chunks, c) for c in range(n_chunks)]
        out_d = out_i = None
        pending = []
        for lo, hi in bounds:
            rec = self.local.search_packed(queries[lo:hi].contiguous(), k, self.id_offset, self.id_stride)     # [q,k,3] int32
            if self.world == 1:
                pending.append((None, rec.reshape(1, hi - lo, k, 3), lo, hi))
                continue
            gathered = torch.empty((self.world * (hi - lo), k * 3), dtype=rec.dtype, device=rec.device)
            # rank-major concatenation == [G][q][k][3] in memory (the layout fire_knn_merge_packed reads)
            work = dist.all_gather_into_tensor(gathered, rec.reshape(hi - lo, k * 3), group=self.group, async_op=True)
            pending.append((work, gathered.view(self.world, hi - lo, k, 3), lo, hi))
        for work, gathered, lo, hi in pending:       # by now every scan is enqueued: chunk c's exchange ran under chunk c+1's scan
            if work is not None:
                work.wait()
            if out_d is None:
                out_d = torch.empty((Q, k), dtype=torch.float32, device=gathered.device)
                out_i = torch.empty((Q, k), dtype=torch.int64, device=gathered.device)
            self.merge_fn(gathered, out_d[lo:hi], out_i[lo:hi])         # row slices are contiguous: merged in place
        return out_d, out_i


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
