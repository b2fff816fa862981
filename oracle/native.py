"""ctypes wrapper around oracle/_build/libfire_oracle.so (see fire_oracle.c for what it restates).

ORACLE - test infrastructure only; the product never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfire_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("fire_oracle.c", "hnsw_oracle.c", "warp_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.fire_oracle_normalize.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        L.fire_oracle_normalize.restype = None
        L.fire_oracle_bf_knn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.fire_oracle_bf_knn.restype = ctypes.c_int
        L.fire_oracle_bf_knn_blocked.argtypes = L.fire_oracle_bf_knn.argtypes
        L.fire_oracle_bf_knn_blocked.restype = ctypes.c_int
        L.fire_oracle_resize_area_u8c3.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.fire_oracle_resize_area_u8c3.restype = ctypes.c_int
        L.fire_oracle_crop_preprocess.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_long] + \
            [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
        L.fire_oracle_crop_preprocess.restype = ctypes.c_int
        L.fire_oracle_get_affine.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.fire_oracle_get_affine.restype = ctypes.c_int
        L.fire_oracle_warp_affine_u8c3.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.fire_oracle_warp_affine_u8c3.restype = ctypes.c_int
        L.fire_hnsw_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        L.fire_hnsw_create.restype = ctypes.c_void_p
        L.fire_hnsw_destroy.argtypes = [ctypes.c_void_p]
        L.fire_hnsw_destroy.restype = None
        L.fire_hnsw_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.fire_hnsw_add.restype = ctypes.c_int
        L.fire_hnsw_count.argtypes = [ctypes.c_void_p]
        L.fire_hnsw_count.restype = ctypes.c_int
        L.fire_hnsw_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]
        L.fire_hnsw_query.restype = ctypes.c_int
        _lib = L
    return _lib


def normalize(x: np.ndarray) -> np.ndarray:
    """hnswlib cosine-space normalisation of every row (bindings.cpp normalize_vector)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    x2 = x.reshape(-1, x.shape[-1])
    out = np.empty_like(x2)
    lib().fire_oracle_normalize(x2.ctypes.data, out.ctypes.data, x2.shape[0], x2.shape[1])
    return out.reshape(x.shape)


class BFIndexOracle:
    """hnswlib.BFIndex(space='cosine') restated: add_items / knn_query / get_current_count."""

    def __init__(self, dim: int):
        self.dim = dim
        self.rows = np.zeros((0, dim), dtype=np.float32)
        self.labels = np.zeros((0,), dtype=np.uint64)

    def add_items(self, data, ids=None):
        data = np.asarray(data, dtype=np.float32).reshape(-1, self.dim)
        if ids is None:
            ids = np.arange(len(self.labels), len(self.labels) + len(data), dtype=np.uint64)
        ids = np.asarray(ids, dtype=np.uint64).reshape(-1)
        self.rows = np.concatenate([self.rows, normalize(data)], 0)
        self.labels = np.concatenate([self.labels, ids], 0)

    def get_current_count(self) -> int:
        return len(self.labels)

    def knn_query(self, data, k: int = 1, num_threads: int = 1, blocked: bool = False):
        """blocked=True: same arithmetic per (query, row) pair, gallery walked once per block of 8 queries (bit-identical
        output, far less DRAM traffic: what the BASELINE-sized parity checks use)."""
        q = normalize(np.asarray(data, dtype=np.float32).reshape(-1, self.dim))
        if k > len(self.labels):
            raise RuntimeError("Cannot return the results in a contiguous 2D array. Probably ef or M is too small")
        labels = np.empty((q.shape[0], k), dtype=np.uint64)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        rows = np.ascontiguousarray(self.rows)
        fn = lib().fire_oracle_bf_knn_blocked if blocked else lib().fire_oracle_bf_knn
        rc = fn(rows.ctypes.data, self.labels.ctypes.data, rows.shape[0], self.dim, q.ctypes.data,
                                      q.shape[0], k, labels.ctypes.data, dists.ctypes.data, num_threads)
        assert rc == 0
        return labels, dists


def resize_area(img: np.ndarray, dh: int = 160, dw: int = 160) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA) for uint8 HxWx3, restated."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    out = np.empty((dh, dw, 3), dtype=np.uint8)
    base = img.ctypes.data
    rc = lib().fire_oracle_resize_area_u8c3(base, img.shape[0], img.shape[1], img.strides[0], out.ctypes.data, dh, dw)
    assert rc == 0 and img.strides[1] == 3 and img.strides[2] == 1
    return out


def crop_preprocess(frame: np.ndarray, box) -> tuple:
    """face_recognition.py:412-420 clamp + slice, then encoder.py:19-27.  Returns (status, u8[160,160,3], f32)."""
    frame = np.ascontiguousarray(frame)
    u8 = np.zeros((160, 160, 3), dtype=np.uint8)
    f32 = np.zeros((160, 160, 3), dtype=np.float32)
    x, y, w, h = (int(v) for v in box)
    rc = lib().fire_oracle_crop_preprocess(frame.ctypes.data, frame.shape[0], frame.shape[1], frame.strides[0],
                                           x, y, w, h, u8.ctypes.data, f32.ctypes.data)
    return rc, u8, f32


def get_affine_transform(src_pts, dst_pts) -> np.ndarray:
    """cv2.getAffineTransform restated (oracle/warp_oracle.c): float32 [3,2] point pairs -> float64 [2,3]."""
    s = np.ascontiguousarray(src_pts, dtype=np.float32).reshape(3, 2)
    d = np.ascontiguousarray(dst_pts, dtype=np.float32).reshape(3, 2)
    M = np.zeros(6, dtype=np.float64)
    lib().fire_oracle_get_affine(s.ctypes.data, d.ctypes.data, M.ctypes.data)
    return M.reshape(2, 3)


def warp_affine(image: np.ndarray, M: np.ndarray, dsize=(160, 160)) -> np.ndarray:
    """cv2.warpAffine(image, M, dsize) restated: INTER_LINEAR, BORDER_CONSTANT 0, uint8 HWC3."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    M = np.ascontiguousarray(M, dtype=np.float64).reshape(6)
    out = np.zeros((dsize[1], dsize[0], 3), dtype=np.uint8)
    lib().fire_oracle_warp_affine_u8c3(image.ctypes.data, image.shape[0], image.shape[1], image.strides[0], M.ctypes.data,
                                       out.ctypes.data, dsize[1], dsize[0])
    return out


class HnswOracle:
    """hnswlib.Index(space='cosine') restated (oracle/hnsw_oracle.c): the APPROXIMATE index the reference queries
    (modules/hnsw_manager.py:20,29-30: max_elements 100000, ef_construction 200, M 16, ef 200).  Used only to report
    recall@k of the reference's index against the exact result."""

    def __init__(self, dim: int, max_elements: int = 100000, M: int = 16, ef_construction: int = 200, seed: int = 100):
        self.dim, self.ef = dim, 10                      # hnswlib's default ef until set_ef() is called
        self._h = lib().fire_hnsw_create(dim, max_elements, M, ef_construction, seed)

    def set_ef(self, ef: int):
        self.ef = int(ef)

    def add_items(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, self.dim)
        for r in rows:                                    # the reference adds one row per call (hnsw_manager.py:127,137)
            if lib().fire_hnsw_add(self._h, r.ctypes.data) < 0:
                raise RuntimeError("index full")

    def get_current_count(self) -> int:
        return int(lib().fire_hnsw_count(self._h))

    def knn_query(self, q: np.ndarray, k: int = 1):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        labels = np.empty((q.shape[0], k), dtype=np.int64)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        rc = lib().fire_hnsw_query(self._h, q.ctypes.data, q.shape[0], k, self.ef, labels.ctypes.data, dists.ctypes.data)
        if rc != 0:
            raise RuntimeError("Cannot return the results in a contiguous 2D array. Probably ef or M is too small")
        return labels.astype(np.uint64), dists

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().fire_hnsw_destroy(self._h)
                self._h = None
        except Exception:
            pass
