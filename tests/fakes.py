"""Test doubles (CPU): an oracle-backed stand-in for fire_b200.engine.KnnIndex so that HOST logic
(HNSWManager, hnswlib_compat.Index, ShardedGallery plumbing) can be exercised without a GPU."""
import numpy as np

from oracle import native


class FakeKnnIndex:
    def __init__(self, dim, capacity=100000, device=0):
        self.dim, self._cap = dim, capacity
        self._ora = native.BFIndexOracle(dim)

    count = property(lambda self: self._ora.get_current_count())
    capacity = property(lambda self: self._cap)

    def close(self):
        pass

    def reset(self):
        self._ora = native.BFIndexOracle(self.dim)

    def add(self, rows):
        rows = np.asarray(rows, dtype=np.float32).reshape(-1, self.dim)
        if self.count + len(rows) > self._cap:
            raise RuntimeError("capacity")
        self._ora.add_items(rows)

    def rows(self, first=0, n=None):
        n = self.count - first if n is None else n
        return self._ora.rows[first:first + n].copy()

    def search(self, queries, k=1, id_offset=0):
        import torch
        was_tensor = torch.is_tensor(queries)
        q = queries.numpy() if was_tensor else np.asarray(queries, dtype=np.float32)
        labels, dist = self._ora.knn_query(q.reshape(-1, self.dim), k)
        ids = labels.astype(np.int64) + id_offset
        return (torch.from_numpy(dist), torch.from_numpy(ids)) if was_tensor else (dist, ids)


    def search_rows(self, first, n, k, id_offset=0):
        labels, dist = self._ora.knn_query(self._ora.rows[first:first + n], k)     # stored rows are unit-norm already
        return dist, labels.astype(np.int64) + id_offset

    def search_packed(self, queries, k, id_offset=0, id_stride=1):
        """What fire_knn_search_packed returns: int32 [Q,k,3] = {distance bits, id low, id high}; k may exceed count."""
        import torch
        q = (queries.numpy() if torch.is_tensor(queries) else np.asarray(queries, dtype=np.float32)).reshape(-1, self.dim)
        kk = min(k, self.count)
        dist = np.full((q.shape[0], k), np.finfo(np.float32).max, np.float32)
        ids = np.full((q.shape[0], k), -1, np.int64)
        if kk > 0:
            labels, d = self._ora.knn_query(q, kk)
            dist[:, :kk] = d
            ids[:, :kk] = labels.astype(np.int64) * id_stride + id_offset
        return torch.from_numpy(pack_records(dist, ids))


def pack_records(dist, ids):
    rec = np.empty(dist.shape + (3,), np.int32)
    rec[..., 0] = dist.view(np.int32)
    u = ids.astype(np.int64).view(np.uint64)
    rec[..., 1] = (u & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    rec[..., 2] = (u >> np.uint64(32)).astype(np.uint32).view(np.int32)
    return rec


def unpack_records(rec):
    rec = np.ascontiguousarray(rec)
    dist = rec[..., 0].copy().view(np.float32)
    u = rec[..., 1].astype(np.uint32).astype(np.uint64) | (rec[..., 2].astype(np.uint32).astype(np.uint64) << np.uint64(32))
    return dist, u.view(np.int64)


def numpy_merge_packed(records, out_d=None, out_i=None):
    """fire_knn_merge_packed restated on the host: records int32 [G,Q,k,3] -> (dist [Q,k], ids [Q,k]); padding sorts last."""
    import torch
    d, i = unpack_records(records.numpy() if torch.is_tensor(records) else records)
    key = np.where(i < 0, np.iinfo(np.int64).max, i)                       # (FLT_MAX, -1) padding loses every tie
    G, Q, k = d.shape
    od, oi = np.empty((Q, k), np.float32), np.empty((Q, k), np.int64)
    for q in range(Q):
        pairs = sorted(zip(d[:, q].ravel().tolist(), key[:, q].ravel().tolist(), i[:, q].ravel().tolist()))[:k]
        od[q] = [p[0] for p in pairs]
        oi[q] = [p[2] for p in pairs]
    if out_d is not None:
        out_d.copy_(torch.from_numpy(od)); out_i.copy_(torch.from_numpy(oi))
        return out_d, out_i
    return torch.from_numpy(od), torch.from_numpy(oi)


def numpy_merge(gd, gi):
    """Reference merge of per-shard lists [G,Q,k] by (distance asc, id asc) - what fire_knn_merge computes."""
    import torch
    d = gd.numpy() if torch.is_tensor(gd) else gd
    i = gi.numpy() if torch.is_tensor(gi) else gi
    G, Q, k = d.shape
    od, oi = np.empty((Q, k), np.float32), np.empty((Q, k), np.int64)
    for q in range(Q):
        pairs = sorted(zip(d[:, q].ravel().tolist(), i[:, q].ravel().tolist()))[:k]
        od[q] = [p[0] for p in pairs]
        oi[q] = [p[1] for p in pairs]
    return torch.from_numpy(od), torch.from_numpy(oi)
