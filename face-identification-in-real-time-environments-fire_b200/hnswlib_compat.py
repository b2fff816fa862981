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

    # ---- persistence: own flat format is written; hnswlib 0.8.0 `save_index` files are READ -----
    def save_index(self, path: str):
        n = self.get_current_count()
        rows = self._idx.rows() if n else np.zeros((0, self.dim), np.float32)
        with open(path, "wb") as f:
            f.write(_MAGIC)
            f.write(struct.pack("<qqqq", self.dim, n, self.max_elements, self.ef))
            f.write(self._labels.astype("<u8").tobytes())
            f.write(rows.astype("<f4").tobytes())

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
