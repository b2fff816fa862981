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
