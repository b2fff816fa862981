"""``hnswlib``-shaped objects backed by the B200 exact cosine index (fire_b200.engine.KnnIndex).

The reference drives hnswlib 0.8.0 through `hnswlib.Index(space='cosine', dim)` and the calls
init_index / set_ef / add_items / knn_query / get_current_count / save_index / load_index
(modules/hnsw_manager.py:20,29-30,43,62,127,137,147,237).  `Index` and `BFIndex` here accept the
same calls.  The search is EXACT (it equals hnswlib.BFIndex) - the graph parameters
(ef, M, ef_construction) are accepted and recorded but do not change results.

Only the cosine space is implemented (the only one the reference uses).
"""
from __future__ import annotations

import struct

import numpy as np

from . import engine as _engine

_MAGIC = b"FIREKNN1"


class Index:
    def __init__(self, space: str = "cosine", dim: int = 128):
        if space != "cosine":
            raise NotImplementedError(f"fire_b200 implements the 'cosine' space only, not {space!r}")
        self.space = space
        self.dim = int(dim)
        self.ef = 10
        self.M = 16
        self.ef_construction = 200
        self.max_elements = 0
        self._idx = None
        self._labels = np.zeros((0,), dtype=np.uint64)
        self._identity = True          # labels[i] == i for every row (FIRE's usage)

    # ---- lifecycle ---------------------------------------------------------------------------
    def init_index(self, max_elements: int, ef_construction: int = 200, M: int = 16, random_seed: int = 100,
                   allow_replace_deleted: bool = False):
        self.max_elements = int(max_elements)
        self.ef_construction, self.M = int(ef_construction), int(M)
        if self._idx is not None:
            self._idx.close()
        self._idx = _engine.KnnIndex(self.dim, capacity=self.max_elements)
        self._labels = np.zeros((0,), dtype=np.uint64)
        self._identity = True

    def set_ef(self, ef: int):
        self.ef = int(ef)

    def set_num_threads(self, n: int):
        pass

    def get_current_count(self) -> int:
        return 0 if self._idx is None else self._idx.count

    def get_max_elements(self) -> int:
        return self.max_elements

    def get_ids_list(self):
        return [int(v) for v in self._labels]

    def resize_index(self, new_size: int):
        rows = self._idx.rows() if self._idx is not None and self._idx.count else np.zeros((0, self.dim), np.float32)
        labels = self._labels
        self.init_index(new_size, self.ef_construction, self.M)
        if len(rows):
            self._idx.add(rows)          # rows are already unit-norm; re-normalising is the identity up to 1 ulp
            self._labels = labels
            self._identity = bool(np.array_equal(labels, np.arange(len(labels), dtype=np.uint64)))

    # ---- data --------------------------------------------------------------------------------
    def add_items(self, data, ids=None, num_threads: int = -1, replace_deleted: bool = False):
        if self._idx is None:
            raise RuntimeError("Index not initialized: call init_index first")
        data = np.asarray(data, dtype=np.float32)
        if data.ndim == 1:
            data = data[None, :]
        if data.ndim != 2 or data.shape[1] != self.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")
        n = data.shape[0]
        start = len(self._labels)
        if ids is None:
            ids_arr = np.arange(start, start + n, dtype=np.uint64)
        else:
            ids_arr = np.asarray(ids, dtype=np.uint64).reshape(-1)
            if ids_arr.shape[0] != n:
                raise RuntimeError("Wrong dimensionality of the labels")
        if start + n > self.max_elements:
            raise RuntimeError("The number of elements exceeds the specified limit")
        self._idx.add(data)
        self._labels = np.concatenate([self._labels, ids_arr])
        if self._identity and not np.array_equal(ids_arr, np.arange(start, start + n, dtype=np.uint64)):
            self._identity = False

    def get_items(self, ids=None, return_type: str = "numpy"):
        if ids is None:
            return self._idx.rows()
        lut = {int(l): i for i, l in enumerate(self._labels)}
        rows = self._idx.rows()
        out = np.stack([rows[lut[int(i)]] for i in ids]) if len(ids) else np.zeros((0, self.dim), np.float32)
        return out if return_type == "numpy" else out.tolist()

    def knn_query(self, data, k: int = 1, num_threads: int = -1, filter=None):
        """-> (labels uint64 [Q,k], distances float32 [Q,k]), rows ascending by (distance, label)."""
        if filter is not None:
            raise NotImplementedError("label filters are not supported")
        data = np.asarray(data, dtype=np.float32)
        if data.ndim == 1:
            data = data[None, :]
        if data.ndim != 2 or data.shape[1] != self.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")
        if self._idx is None or k > self._idx.count:
            raise RuntimeError("Cannot return the results in a contiguous 2D array. Probably ef or M is too small")
        dist, rows = self._idx.search(data, k)
        labels = rows.astype(np.uint64) if self._identity else self._labels[rows]
        return labels, dist

    def knn_query_rows(self, first: int, n: int, k: int = 1):
        """knn_query with the STORED rows [first, first+n) as the queries (additive; fire_knn_search_rows): what
        knn_query(original vector of row i) returns for every i in the range, in one GPU pass."""
        if self._idx is None or k > self._idx.count:
            raise RuntimeError("Cannot return the results in a contiguous 2D array. Probably ef or M is too small")
        dist, rows = self._idx.search_rows(first, n, k)
        dist = dist.cpu().numpy() if hasattr(dist, "cpu") else np.asarray(dist)
        rows = rows.cpu().numpy() if hasattr(rows, "cpu") else np.asarray(rows)
        labels = rows.astype(np.uint64) if self._identity else self._labels[rows]
        return labels, dist

    # ---- persistence: the file is an hnswlib 0.8.0 `save_index` image in BOTH directions, so a storage/
    # tree written here still opens with the reference's hnswlib.load_index (a roll-back does not lose the gallery) ---------
    def save_index(self, path: str):
        n = self.get_current_count()
        rows = self._idx.rows() if n else np.zeros((0, self.dim), np.float32)
        links = self._knn_graph(n, 2 * self.M)
        with open(path, "wb") as f:
            f.write(write_hnswlib_binary(self._labels, rows, links, self.max_elements, self.M, self.ef_construction))

    def _knn_graph(self, n: int, max_links: int) -> np.ndarray:
        """Level-0 links of the exported graph: every element's exact nearest neighbours (the GPU search finds them in a
        few tiled passes).  An exact k-NN graph is what HNSW's construction approximates on its bottom layer."""
        k = min(max_links, n - 1)
        if k <= 0:
            return np.zeros((n, 0), dtype=np.uint32)
        out = np.empty((n, k), dtype=np.uint32)
        for first in range(0, n, 8192):
            m = min(8192, n - first)
            _, nb = self._idx.search_rows(first, m, k + 1)
            nb = nb.cpu().numpy() if hasattr(nb, "cpu") else np.asarray(nb)
            me = np.arange(first, first + m)[:, None]
            keep = nb != me                                        # drop the row itself (an exact duplicate may sit before it)
            keep[keep.sum(1) > k, -1] = False                      # no self hit among k+1 (k+1 duplicates): drop the farthest
            out[first:first + m] = nb[keep].reshape(m, k)
        return out

    def load_index(self, path: str, max_elements: int = 0, allow_replace_deleted: bool = False):
        with open(path, "rb") as f:
            blob = f.read()
        if blob[:8] != _MAGIC:
            labels, rows, saved_max = parse_hnswlib_binary(blob, self.dim)     # an existing FIRE storage/ tree (hnsw_manager.py:43,62)
            n, dim = rows.shape
        else:
            dim, n, saved_max, _ef = struct.unpack("<qqqq", blob[8:40])
            if dim != self.dim:
                raise RuntimeError(f"Index dimensionality {dim} does not match {self.dim}")
            labels = np.frombuffer(blob, dtype="<u8", count=n, offset=40).astype(np.uint64)
            rows = np.frombuffer(blob, dtype="<f4", count=n * dim, offset=40 + 8 * n).reshape(n, dim)
        self.init_index(max(int(max_elements) if max_elements else int(saved_max), n), self.ef_construction, self.M)
        self.ef = 10                       # hnswlib does not persist ef
        if n:
            self._idx.add(np.ascontiguousarray(rows))
            self._labels = labels
            self._identity = bool(np.array_equal(labels, np.arange(n, dtype=np.uint64)))


_HNSW_HEADER = struct.Struct("<6QiI3QdQ")      # HierarchicalNSW::saveIndex, hnswlib 0.8.0 (hnswalg.h)


def write_hnswlib_binary(labels: np.ndarray, rows: np.ndarray, links: np.ndarray, max_elements: int, M: int = 16,
                         ef_construction: int = 200) -> bytes:
    """The inverse of `parse_hnswlib_binary`: an hnswlib 0.8.0 `save_index` image (layout documented there) of a one-level
    graph - every element on level 0, entry point 0, level-0 neighbour lists = `links` (uint32 [n, <= 2M] internal ids),
    no upper-level lists (one zero uint32 per element)."""
    n, dim = rows.shape
    maxM0 = 2 * M
    assert links.shape[0] == n and links.shape[1] <= maxM0
    data_off = maxM0 * 4 + 4
    label_off = data_off + dim * 4
    per_el = label_off + 8
    rec = np.zeros((n, per_el), dtype=np.uint8)
    if n:
        head = np.zeros((n, 1 + maxM0), dtype="<u4")
        head[:, 0] = links.shape[1]
        head[:, 1:1 + links.shape[1]] = links
        rec[:, :data_off] = head.view(np.uint8).reshape(n, data_off)
        rec[:, data_off:label_off] = np.ascontiguousarray(rows, dtype="<f4").view(np.uint8).reshape(n, dim * 4)
        rec[:, label_off:] = np.ascontiguousarray(labels, dtype="<u8").view(np.uint8).reshape(n, 8)
    header = _HNSW_HEADER.pack(0, max(int(max_elements), n), n, per_el, label_off, data_off, 0 if n else -1, 0 if n else 0xFFFFFFFF,
                               M, maxM0, M, 1.0 / np.log(M), ef_construction)
    return header + rec.tobytes() + np.zeros(n, dtype="<u4").tobytes()


def parse_hnswlib_binary(blob: bytes, dim: int):
    """Read the vectors and labels out of a file written by hnswlib 0.8.0 `Index.save_index` (the graph is not needed:
    the B200 search is exact).  Layout restated from hnswalg.h `saveIndex` (the library is not vendored in the
    reference): a 96-byte header
        offsetLevel0, max_elements, cur_element_count, size_data_per_element, label_offset, offsetData (size_t each),
        maxlevel (int), enterpoint (uint), maxM, maxM0, M (size_t), mult (double), ef_construction (size_t)
    then cur_element_count level-0 records of size_data_per_element bytes
        [uint32 link count/flags][maxM0 x uint32 links][dim x float32 vector][uint64 label]
    then the upper-level link lists.  Cosine-space vectors are stored normalised.  Returns (labels uint64 [n],
    rows float32 [n, dim], max_elements); elements carrying hnswlib's delete mark are dropped."""
    if len(blob) < _HNSW_HEADER.size:
        raise RuntimeError("Index seems to be corrupted or unsupported")
    (off0, max_el, count, per_el, label_off, data_off, _maxlevel, _enter, maxM, maxM0, M, _mult, _efc) = _HNSW_HEADER.unpack_from(blob, 0)
    sane = (off0 == 0 and data_off == maxM0 * 4 + 4 and label_off == data_off + dim * 4 and per_el == label_off + 8 and
            0 < M <= 4096 and maxM == M and maxM0 == 2 * M and count <= max_el and _HNSW_HEADER.size + count * per_el <= len(blob))
    if not sane:
        raise RuntimeError("Index seems to be corrupted or unsupported (neither a fire_b200 nor an hnswlib 0.8.0 cosine index "
                           f"of dimension {dim})")
    rec = np.frombuffer(blob, dtype=np.uint8, count=count * per_el, offset=_HNSW_HEADER.size).reshape(count, per_el)
    rows = np.ascontiguousarray(rec[:, data_off:label_off]).view("<f4").reshape(count, dim)
    labels = np.ascontiguousarray(rec[:, label_off:label_off + 8]).view("<u8").reshape(count)
    flags = np.ascontiguousarray(rec[:, 0:4]).view("<u4").reshape(count)
    live = ((flags >> 16) & 1) == 0                   # DELETE_MARK lives in the third byte of the level-0 link header
    labels, rows = labels[live].astype(np.uint64), rows[live].astype(np.float32)
    order = np.argsort(labels, kind="stable")         # FIRE's ids are its running counter: row index == label
    return labels[order], np.ascontiguousarray(rows[order]), int(max_el)


class BFIndex(Index):
    """hnswlib.BFIndex: same exact search; init_index takes only max_elements."""

    def init_index(self, max_elements: int, *args, **kwargs):
        super().init_index(max_elements)
