"""ctypes binding of libfire_b200.so (C ABI: include/fire_b200.h).

There is no fallback of any kind: if the shared library is missing this module raises, and if no
sm_100 device is usable every compute entry point returns an error that `check()` turns into
`FireError`.  Build the library with ``python -m fire_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libfire_b200.so")

FIRE_OK = 0
PRE_REFERENCE, PRE_NORTHSTAR, PRE_FLAG_SWAP_RB = 0, 1, 16


class FireError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "fire_init": (C.c_int, [C.c_int]),
    "fire_last_error": (C.c_char_p, []),
    "fire_version": (C.c_char_p, []),
    "fire_launch_count": (C.c_uint64, []),
    "fire_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fire_roi_meta_bytes": (C.c_size_t, [C.c_int]),
    "fire_pack_rois_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_size_t), C.c_int]),
    "fire_upload_rois_dma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.c_void_p]),
    "fire_align_warp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "fire_ingest_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "fire_facenet_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "fire_facenet_destroy": (C.c_int, [C.c_void_p]),
    "fire_facenet_dim": (C.c_int, [C.c_void_p]),
    "fire_facenet_workspace": (C.c_size_t, [C.c_void_p, C.c_int]),
    "fire_facenet_flops": (C.c_double, [C.c_void_p]),
    "fire_facenet_num_ops": (C.c_int, [C.c_void_p]),
    "fire_facenet_num_launches": (C.c_int, [C.c_void_p]),
    "fire_facenet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "fire_facenet_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p]),
    "fire_facenet_read_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_size_t]),
    "fire_knn_create": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "fire_knn_destroy": (C.c_int, [C.c_void_p]),
    "fire_knn_reset": (C.c_int, [C.c_void_p]),
    "fire_knn_count": (C.c_size_t, [C.c_void_p]),
    "fire_knn_capacity": (C.c_size_t, [C.c_void_p]),
    "fire_knn_dim": (C.c_int, [C.c_void_p]),
    "fire_knn_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fire_knn_add_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fire_knn_get_rows_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "fire_knn_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "fire_knn_search_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "fire_knn_search_rows": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fire_knn_search_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "fire_knn_merge_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fire_knn_stats_ex": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "fire_knn_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "fire_knn_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fire_knn_set_margin": (C.c_int, [C.c_void_p, C.c_float]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded library (loads on first use; raises FireError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FireError(f"{LIB_PATH} is missing: build it with `python -m fire_b200.build` "
                            "(fire_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != FIRE_OK:
        raise FireError(f"libfire_b200 error {rc}: {lib().fire_last_error().decode(errors='replace')}")


def init(device: int = 0) -> None:
    """cudaSetDevice + capability check (sm_100 required).  Raises FireError on a GPU-less host.  Always makes `device`
    current: a handle lives on the device that is current when it is created (every later call on the handle switches
    back to that device by itself)."""
    check(lib().fire_init(device))


def launch_count() -> int:
    return int(lib().fire_launch_count())


def stream_ptr(stream=None) -> int:
    """cudaStream_t of a torch stream (None -> torch's current stream) as an integer for ctypes."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)
