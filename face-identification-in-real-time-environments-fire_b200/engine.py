"""Device-level Python API over libfire_b200.so: torch tensors own the HBM, the kernels are ours.

These classes are what the drop-in surfaces (encoder.py, hnsw_manager.py, preprocess.py) and
bench.py are built from.  torch is used for device memory, streams and (in dist.py) NCCL only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import FireError, check
from .netplan import IN_HW, Plan

NET_HW, NET_C = IN_HW // 2, 16          # network input: space-to-depth fp16 [B, 80, 80, 16] (netplan.Plan, csrc/preprocess.cu)
from . import weights as W


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise FireError("fire_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def _ptr(t) -> int:
    return int(t.data_ptr())


class FaceNetEngine:
    """K2: the Inception-ResNet-v1 conv stack (replaces the onnxruntime session, facenet_gpu.py:72,127)."""

    def __init__(self, D: int = 512, tensors: Optional[dict] = None, device: int = 0, fuse_siblings: bool = True,
                 seed: int = 1234, pitched: bool = True, reuse_buffers: bool = True, pair_stem: bool = True):
        torch = _torch()
        _lib.init(device)
        self.device = torch.device("cuda", device)
        self.D = D
        self.plan = Plan(D, fuse_siblings=fuse_siblings, pitched=pitched, reuse_buffers=reuse_buffers, pair_stem=pair_stem)   # reuse_buffers=False: debug layout, every activation keeps its own memory
        if tensors is None:
            tensors = W.synthetic_weights(D, seed)
        self.blob = W.pack(self.plan, tensors)
        h = C.c_void_p()
        buf = C.create_string_buffer(self.blob, len(self.blob))
        check(_lib.lib().fire_facenet_create(C.addressof(buf), len(self.blob), C.byref(h)))
        self._h = h
        self._ws = None
        self._ws_B = 0
        self.flops_per_image = float(_lib.lib().fire_facenet_flops(self._h))
        self.num_ops = int(_lib.lib().fire_facenet_num_ops(self._h))
        self.num_launches = int(_lib.lib().fire_facenet_num_launches(self._h))   # fused chains count once

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().fire_facenet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, B: int):
        torch = _torch()
        need = int(_lib.lib().fire_facenet_workspace(self._h, B))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        off = (-self._ws.data_ptr()) % 256
        return self._ws.data_ptr() + off, self._ws.numel() - off

    def forward(self, x_f16, want_l2: bool = True, out_raw=None, out_l2=None):
        """x_f16: cuda fp16 [B,80,80,16] space-to-depth pixel-scale network input (what preprocess_boxes / ingest_unit_f32
        return; pixels_to_network_input() builds it from a plain image batch) -> (raw [B,D] f32, l2 [B,D] f32 or None)."""
        torch = _torch()
        assert x_f16.is_cuda and x_f16.dtype == torch.float16 and x_f16.is_contiguous()
        B = x_f16.shape[0]
        assert tuple(x_f16.shape[1:]) == (NET_HW, NET_HW, NET_C), x_f16.shape
        raw = out_raw if out_raw is not None else torch.empty(B, self.D, dtype=torch.float32, device=self.device)
        l2 = (out_l2 if out_l2 is not None else torch.empty(B, self.D, dtype=torch.float32, device=self.device)) if want_l2 else None
        ws_ptr, ws_bytes = self._workspace(B)
        check(_lib.lib().fire_facenet_forward(self._h, _ptr(x_f16), B, _ptr(raw), _ptr(l2) if l2 is not None else None,
                                              ws_ptr, ws_bytes, _lib.stream_ptr()))
        return raw, l2

    def ingest_unit_f32(self, x_f32):
        """cuda float32 [B,160,160,3] in the reference's [0,1] scale -> fp16 [B,80,80,16] network input."""
        torch = _torch()
        assert x_f32.is_cuda and x_f32.dtype == torch.float32 and x_f32.is_contiguous()
        B = x_f32.shape[0]
        out = torch.empty(B, NET_HW, NET_HW, NET_C, dtype=torch.float16, device=self.device)
        check(_lib.lib().fire_ingest_f32(_ptr(x_f32), B, _ptr(out), _lib.stream_ptr()))
        return out

    def encode_unit_f32(self, x_f32, want_l2: bool = True):
        return self.forward(self.ingest_unit_f32(x_f32), want_l2=want_l2)

    def profile(self, x_f16) -> Tuple[np.ndarray, np.ndarray]:
        """Per-op device milliseconds and algorithmic FLOP for one forward of this batch."""
        B = x_f16.shape[0]
        ms = np.zeros(self.num_ops, dtype=np.float32)
        fl = np.zeros(self.num_ops, dtype=np.float64)
        ws_ptr, ws_bytes = self._workspace(B)
        check(_lib.lib().fire_facenet_profile(self._h, _ptr(x_f16), B, ws_ptr, ws_bytes, ms.ctypes.data, fl.ctypes.data,
                                              self.num_ops, _lib.stream_ptr()))
        return ms, fl

    def read_buffer(self, buf: int, x_f16) -> np.ndarray:
        """Debug: activation buffer `buf` of the last forward on x_f16, as float32 [B,H,W,C]."""
        B = x_f16.shape[0]
        b = self.plan.bufs[buf]
        out = np.empty((B, b.H, b.Wp or b.W, b.C), dtype=np.float16)
        ws_ptr, _ = self._workspace(B)
        _torch().cuda.synchronize()
        check(_lib.lib().fire_facenet_read_buffer(self._h, buf, B, _ptr(x_f16), ws_ptr, out.ctypes.data, out.nbytes))
        return out[:, :, :b.W, :].astype(np.float32)          # drop the row-pitch padding of pitched buffers



def pixels_to_network_input(x):
    """torch [B,160,160,C>=3] pixel-scale values (any float/int dtype) -> fp16 [B,80,80,16] space-to-depth network input:
    channel (dy*2+dx)*3+c of position (Y,X) = pixel (2Y+dy, 2X+dx), channel c.  Plain torch (tests / tools only; the
    product path gets this layout straight out of fire_preprocess / fire_ingest_f32)."""
    torch = _torch()
    B = x.shape[0]
    y = x[..., :3].reshape(B, NET_HW, 2, NET_HW, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, NET_HW, NET_HW, 12)
    out = torch.zeros(B, NET_HW, NET_HW, NET_C, dtype=torch.float16, device=x.device)
    out[..., :12] = y.to(torch.float16)
    return out


def network_input_to_pixels(f16):
    """Inverse of pixels_to_network_input: fp16 [B,80,80,16] -> float32 [B,160,160,3] (also returns the 4 padding channels)."""
    B = f16.shape[0]
    y = f16[..., :12].float().reshape(B, NET_HW, NET_HW, 2, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, IN_HW, IN_HW, 3)
    return y, f16[..., 12:].float()


class CropEncodePipeline:
    """Streaming encode of fixed-size batches of 160x160 uint8 crops held in PINNED host memory.

    The reference encodes one face per ``session.run`` (modules/face_recognition.py:404-486); this is the batched
    counterpart a caller that collects crops would use.  ``submit()`` enqueues, without synchronising,
      host -> device copy (its own stream)  ->  K1 crop/resize/normalise  ->  K2 FaceNet  ->  device -> host copy
    with ``depth`` staging buffers, so the copy of batch i+1 overlaps the kernels of batch i.
    ``result(ticket)`` waits for that batch only and returns the pinned float32 [B, D] embeddings (L2-normalised
    like modules/face_recognition.py:225-229 when ``normalize``)."""

    def __init__(self, engine: "FaceNetEngine", batch: int, depth: int = 2, normalize: bool = True,
                 mode: int = _lib.PRE_REFERENCE):
        torch = _torch()
        self.eng, self.batch, self.depth, self.normalize, self.mode = engine, batch, max(2, depth), normalize, mode
        dev = engine.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty(batch, IN_HW, IN_HW, 3, dtype=torch.uint8, device=dev) for _ in range(self.depth)]
        self.raw = [torch.empty(batch, engine.D, dtype=torch.float32, device=dev) for _ in range(self.depth)]
        self.l2 = [torch.empty(batch, engine.D, dtype=torch.float32, device=dev) for _ in range(self.depth)]
        self.host = [torch.empty(batch, engine.D, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(self.depth)]      # H2D of slot landed
        self.ev_free = [torch.cuda.Event() for _ in range(self.depth)]    # kernels reading the slot's staging buffer are done
        self.ev_out = [torch.cuda.Event() for _ in range(self.depth)]     # D2H of slot landed
        self.boxes = torch.tensor([[0, 0, IN_HW, IN_HW]] * batch, dtype=torch.int32, device=dev)
        self.box_frame = torch.arange(batch, dtype=torch.int32, device=dev)
        self.desc = torch.tensor([[i * IN_HW * IN_HW * 3, IN_HW, IN_HW, IN_HW * 3] for i in range(batch)], dtype=torch.int64, device=dev)
        self.n = 0
        self.h2d_bytes = batch * IN_HW * IN_HW * 3
        self.d2h_bytes = batch * engine.D * 4

    def submit(self, crops_pinned_u8) -> int:
        torch = _torch()
        assert crops_pinned_u8.dtype == torch.uint8 and tuple(crops_pinned_u8.shape) == (self.batch, IN_HW, IN_HW, 3)
        slot = self.n % self.depth
        main = torch.cuda.current_stream()
        if self.n >= self.depth:
            self.copy_stream.wait_event(self.ev_free[slot])            # do not overwrite a staging buffer still being read
            self.ev_out[slot].synchronize()                            # the caller may still be reading host[slot]? it was handed out `depth` submits ago
        with torch.cuda.stream(self.copy_stream):
            self.stage[slot].copy_(crops_pinned_u8, non_blocking=True)
            self.ev_in[slot].record(self.copy_stream)
        main.wait_event(self.ev_in[slot])
        f16, _, _ = preprocess_boxes(self.stage[slot], self.desc, self.boxes, self.box_frame, self.mode, True, False)
        self.ev_free[slot].record(main)
        self.eng.forward(f16, want_l2=self.normalize, out_raw=self.raw[slot], out_l2=self.l2[slot] if self.normalize else None)
        self.host[slot].copy_(self.l2[slot] if self.normalize else self.raw[slot], non_blocking=True)
        self.ev_out[slot].record(main)
        self.n += 1
        return self.n - 1

    def result(self, ticket: int):
        assert self.n - self.depth <= ticket < self.n, "result() must be read before `depth` later submits reuse the slot"
        slot = ticket % self.depth
        self.ev_out[slot].synchronize()
        return self.host[slot]

    def drain(self):
        _torch().cuda.current_stream().synchronize()


class KnnIndex:
    """K3: exact cosine top-k over a row-major gallery shard (replaces hnswlib, hnsw_manager.py:20-31,127-149)."""

    def __init__(self, dim: int, capacity: int = 100000, device: int = 0):
        torch = _torch()
        _lib.init(device)
        self.device = torch.device("cuda", device)
        self.dim = dim
        h = C.c_void_p()
        check(_lib.lib().fire_knn_create(dim, capacity, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().fire_knn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def count(self) -> int:
        return int(_lib.lib().fire_knn_count(self._h))

    @property
    def capacity(self) -> int:
        return int(_lib.lib().fire_knn_capacity(self._h))

    def reset(self):
        check(_lib.lib().fire_knn_reset(self._h))

    def add(self, rows):
        """rows: numpy [n,D] float32 (host path) or cuda float32 tensor (device path, async on the current stream)."""
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, self.dim)
            check(_lib.lib().fire_knn_add_host(self._h, rows.ctypes.data, rows.shape[0]))
        else:
            torch = _torch()
            assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and rows.shape[-1] == self.dim
            check(_lib.lib().fire_knn_add(self._h, _ptr(rows), rows.numel() // self.dim, _lib.stream_ptr()))

    def rows(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.count - first if n is None else n
        out = np.empty((n, self.dim), dtype=np.float32)
        check(_lib.lib().fire_knn_get_rows_host(self._h, first, n, out.ctypes.data))
        return out

    def search(self, queries, k: int = 1, id_offset: int = 0, out_dist=None, out_ids=None):
        """Exact top-k.  numpy in -> numpy (dist f32 [Q,k], ids int64 [Q,k]); cuda tensor in -> cuda tensors (async)."""
        if isinstance(queries, np.ndarray):
            q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
            dist = np.empty((q.shape[0], k), dtype=np.float32)
            ids = np.empty((q.shape[0], k), dtype=np.int64)
            check(_lib.lib().fire_knn_search_host(self._h, q.ctypes.data, q.shape[0], k, id_offset, dist.ctypes.data,
                                                  ids.ctypes.data))
            return dist, ids
        torch = _torch()
        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        Q = queries.numel() // self.dim
        dist = out_dist if out_dist is not None else torch.empty(Q, k, dtype=torch.float32, device=self.device)
        ids = out_ids if out_ids is not None else torch.empty(Q, k, dtype=torch.int64, device=self.device)
        check(_lib.lib().fire_knn_search(self._h, _ptr(queries), Q, k, id_offset, _ptr(dist), _ptr(ids), _lib.stream_ptr()))
        return dist, ids

    def search_rows(self, first: int, n: int, k: int, id_offset: int = 0, out_dist=None, out_ids=None):
        """Top-k with the STORED rows [first, first+n) as queries (bulk find_similar_embeddings, hnsw_manager.py:227-244).
        Returns cuda tensors (dist f32 [n,k], ids int64 [n,k]); every row's first hit is itself (distance ~0)."""
        torch = _torch()
        dist = out_dist if out_dist is not None else torch.empty(n, k, dtype=torch.float32, device=self.device)
        ids = out_ids if out_ids is not None else torch.empty(n, k, dtype=torch.int64, device=self.device)
        check(_lib.lib().fire_knn_search_rows(self._h, first, n, k, id_offset, _ptr(dist), _ptr(ids), _lib.stream_ptr()))
        return dist, ids

    def search_packed(self, queries, k: int, id_offset: int = 0, id_stride: int = 1, out=None):
        """Shard search for the multi-GPU exchange: cuda float32 queries -> cuda int32 [Q,k,3] records
        {distance bits, id low, id high}; id = row * id_stride + id_offset; k may exceed this shard's row count
        (missing entries = (FLT_MAX, -1))."""
        torch = _torch()
        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        Q = queries.numel() // self.dim
        rec = out if out is not None else torch.empty(Q, k, 3, dtype=torch.int32, device=self.device)
        check(_lib.lib().fire_knn_search_packed(self._h, _ptr(queries), Q, k, id_offset, id_stride, _ptr(rec), _lib.stream_ptr()))
        return rec

    def stats(self) -> Tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        check(_lib.lib().fire_knn_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def stats_ex(self) -> Tuple[int, int, int, int]:
        """(queries, queries whose merged-list proof failed, of those: single-split exact scans, whole-shard exact scans)."""
        v = (C.c_uint64 * 4)()
        check(_lib.lib().fire_knn_stats_ex(self._h, v))
        return tuple(int(x) for x in v)

    def set_margin(self, eps: float):
        check(_lib.lib().fire_knn_set_margin(self._h, eps))


def knn_merge(dists, ids):
    """dists/ids: cuda [G,Q,k] per-shard results (rows ascending) -> global top-k (dist [Q,k], ids [Q,k])."""
    torch = _torch()
    G, Q, k = dists.shape
    assert dists.is_contiguous() and ids.is_contiguous() and ids.dtype == torch.int64 and dists.dtype == torch.float32
    od = torch.empty(Q, k, dtype=torch.float32, device=dists.device)
    oi = torch.empty(Q, k, dtype=torch.int64, device=dists.device)
    check(_lib.lib().fire_knn_merge(_ptr(dists), _ptr(ids), Q, k, G, _ptr(od), _ptr(oi), _lib.stream_ptr()))
    return od, oi


def knn_merge_packed(records, out_dist=None, out_ids=None):
    """records: cuda int32 [G,Q,k,3] (G shards' KnnIndex.search_packed results, e.g. straight out of one all_gather)
    -> global top-k (dist f32 [Q,k], ids int64 [Q,k])."""
    torch = _torch()
    G, Q, k, three = records.shape
    assert three == 3 and records.is_contiguous() and records.dtype == torch.int32
    od = out_dist if out_dist is not None else torch.empty(Q, k, dtype=torch.float32, device=records.device)
    oi = out_ids if out_ids is not None else torch.empty(Q, k, dtype=torch.int64, device=records.device)
    check(_lib.lib().fire_knn_merge_packed(_ptr(records), Q, k, G, _ptr(od), _ptr(oi), _lib.stream_ptr()))
    return od, oi


def preprocess_boxes(frames, frame_desc, boxes, box_frame, mode: int = _lib.PRE_REFERENCE, want_f16: bool = True,
                     want_f32: bool = False):
    """K1.  frames: cuda uint8 (flat or [F,H,W,3]); frame_desc: cuda int64 [F,4] (offset,H,W,stride);
    boxes: cuda int32 [n,4] xywh; box_frame: cuda int32 [n].  Returns (f16 [n,80,80,16] network input | None,
    f32 [n,160,160,3] | None, status int32 [n])."""
    torch = _torch()
    n = boxes.shape[0]
    dev = frames.device
    f16 = torch.empty(n, NET_HW, NET_HW, NET_C, dtype=torch.float16, device=dev) if want_f16 else None
    f32 = torch.empty(n, IN_HW, IN_HW, 3, dtype=torch.float32, device=dev) if want_f32 else None
    status = torch.empty(n, dtype=torch.int32, device=dev)
    check(_lib.lib().fire_preprocess(_ptr(frames), _ptr(frame_desc), frame_desc.shape[0], _ptr(boxes), _ptr(box_frame), n,
                                     mode, _ptr(f16) if f16 is not None else None, _ptr(f32) if f32 is not None else None,
                                     _ptr(status), _lib.stream_ptr()))
    return f16, f32, status


def pack_rois(frames_host, frame_desc_host, boxes_host, box_frame_host, out_host, threads: int = 4) -> int:
    """fire_pack_rois_host: crop rectangles of `boxes_host` (int32 [n,4] xywh, the reference's clamp rule) out of host frames
    (uint8, flat or [F,H,W,3]; frame_desc_host int64 [F,4] = offset, H, W, row stride) into the staging buffer `out_host`
    (uint8, 16-byte aligned, ideally pinned), tables first.  numpy arrays or CPU torch tensors.  Returns the bytes used.
    Pure host work (no GPU needed): staging for the upload, not compute."""
    def addr(a):
        return a.ctypes.data if isinstance(a, np.ndarray) else int(a.data_ptr())

    def nbytes(a):
        return a.nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size()
    n = int(boxes_host.shape[0])
    used = C.c_size_t(0)
    check(_lib.lib().fire_pack_rois_host(addr(frames_host), addr(frame_desc_host), int(frame_desc_host.shape[0]), addr(boxes_host),
                                         addr(box_frame_host), n, addr(out_host), nbytes(out_host), C.byref(used), threads))
    return int(used.value)


def roi_views(buf, n: int):
    """The tables at the head of a packed ROI buffer (host or device tensor): (frame_desc int64 [n,4], boxes int32 [n,4],
    box_frame int32 [n])."""
    torch = _torch() if buf.is_cuda else __import__("torch")
    desc = buf[:32 * n].view(torch.int64).view(n, 4)
    boxes = buf[32 * n:48 * n].view(torch.int32).view(n, 4)
    bframe = buf[48 * n:52 * n].view(torch.int32)
    return desc, boxes, bframe


class RoiStager:
    """Upload only what the crops read (BASELINE configs[4]): per step, the crop rectangles of all boxes are packed on the
    host into a pinned buffer (fire_pack_rois_host) and moved with ONE host->device copy on a copy stream, `depth` slots
    deep so that the upload of step i+1 runs under the kernels of step i.  `submit` returns the device arguments of
    preprocess_boxes; results are bit-identical to uploading the whole frames.  `submit(..., next_args=...)` also starts
    packing the NEXT step's rectangles on a background thread (the C call releases the GIL), so that the host-side
    gather of step i+1 runs while the main thread enqueues the kernels of step i."""

    def __init__(self, max_bytes: int, depth: int = 2, device: int = 0, threads: int = 4, mode: str = "pack", max_boxes: int = 4096):
        """mode "pack": host gather into a pinned buffer + one copy (fire_pack_rois_host); mode "dma": the copy engine gathers,
        one cudaMemcpy2DAsync per rectangle straight out of the PINNED frames (fire_upload_rois_dma) - no host cores needed."""
        torch = _torch()
        assert mode in ("pack", "dma")
        self.device = torch.device("cuda", device)
        self.depth, self.threads, self.n, self.mode = max(2, depth), threads, 0, mode
        host_bytes = max_bytes if mode == "pack" else int(_lib.lib().fire_roi_meta_bytes(max_boxes))
        self.host = [torch.empty(host_bytes, dtype=torch.uint8).pin_memory() for _ in range(self.depth)]
        self.dev = [torch.empty(max_bytes, dtype=torch.uint8, device=self.device) for _ in range(self.depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        self.ev_free = [torch.cuda.Event() for _ in range(self.depth)]
        self.last_bytes = 0
        self._pending = None            # (thread, slot, args, result holder) of a background pack

    def _pack_into(self, slot: int, args, holder: dict):
        try:
            if self.n >= self.depth or holder.get("ahead"):
                self.ev_in[slot].synchronize()                  # the copy out of this pinned slot (depth submits ago) has finished
            holder["used"] = pack_rois(*args, self.host[slot], self.threads)
        except Exception as e:                                  # re-raised by the submit that consumes this pack
            holder["error"] = e

    def submit(self, frames_host, frame_desc_host, boxes_host, box_frame_host, next_args=None):
        """-> (frames, frame_desc, boxes, box_frame) cuda tensors for preprocess_boxes, valid until `depth` later submits;
        call `release()` once the kernels reading them are enqueued.  next_args: the four arguments of the NEXT submit."""
        import threading
        torch = _torch()
        slot = self.n % self.depth
        args = (frames_host, frame_desc_host, boxes_host, box_frame_host)
        n = int(boxes_host.shape[0])
        if self.mode == "dma":
            def addr(a):
                return a.ctypes.data if isinstance(a, np.ndarray) else int(a.data_ptr())
            if self.n >= self.depth:
                self.ev_in[slot].synchronize()                  # the table copy out of this pinned slot has finished
                self.copy_stream.wait_event(self.ev_free[slot])
            used = C.c_size_t(0)
            check(_lib.lib().fire_upload_rois_dma(addr(frames_host), addr(frame_desc_host), int(frame_desc_host.shape[0]), addr(boxes_host),
                                                  addr(box_frame_host), n, _ptr(self.host[slot]), _ptr(self.dev[slot]), self.dev[slot].numel(),
                                                  C.byref(used), int(self.copy_stream.cuda_stream)))
            self.last_bytes = int(used.value)
            self.ev_in[slot].record(self.copy_stream)
            torch.cuda.current_stream().wait_event(self.ev_in[slot])
            self._slot = slot
            self.n += 1
            return (self.dev[slot],) + roi_views(self.dev[slot], n)
        holder = None
        if self._pending is not None:
            th, p_slot, p_args, p_holder = self._pending
            th.join()
            self._pending = None
            if p_slot == slot and all(a is b for a, b in zip(p_args, args)):
                holder = p_holder
        if holder is None:
            holder = {}
            self._pack_into(slot, args, holder)
        if "error" in holder:
            raise holder["error"]
        used = holder["used"]
        self.last_bytes = used
        if self.n >= self.depth:
            self.copy_stream.wait_event(self.ev_free[slot])     # the kernels that read the device slot are done
        with torch.cuda.stream(self.copy_stream):
            self.dev[slot][:used].copy_(self.host[slot][:used], non_blocking=True)
            self.ev_in[slot].record(self.copy_stream)
        torch.cuda.current_stream().wait_event(self.ev_in[slot])
        self._slot = slot
        self.n += 1
        if next_args is not None:
            nslot = self.n % self.depth
            nh = {"ahead": self.n >= self.depth}
            th = threading.Thread(target=self._pack_into, args=(nslot, tuple(next_args), nh), daemon=True)
            th.start()
            self._pending = (th, nslot, tuple(next_args), nh)
        return (self.dev[slot],) + roi_views(self.dev[slot], n)

    def release(self):
        self.ev_free[self._slot].record(_torch().cuda.current_stream())

    def close(self):
        if self._pending is not None:
            self._pending[0].join()
            self._pending = None


def align_warp(frames, frame_desc, matrices, face_frame, swap_rb: bool = True, output: str = "uint8"):
    """fire_align_warp: frames cuda uint8, frame_desc cuda int64 [F,4], matrices cuda float64 [n,6] (forward 2x3, as
    getAffineTransform returns them), face_frame cuda int32 [n].  output "uint8" -> numpy [n,160,160,3]; "device" ->
    cuda fp16 network input [n,80,80,16]; "both" -> (cuda uint8 [n,160,160,3], cuda fp16 [n,80,80,16])."""
    torch = _torch()
    n = matrices.shape[0]
    assert matrices.dtype == torch.float64 and matrices.is_contiguous() and tuple(matrices.shape) == (n, 6)
    dev = frames.device
    u8 = torch.empty(n, IN_HW, IN_HW, 3, dtype=torch.uint8, device=dev) if output in ("uint8", "both") else None
    f16 = torch.empty(n, NET_HW, NET_HW, NET_C, dtype=torch.float16, device=dev) if output in ("device", "both") else None
    check(_lib.lib().fire_align_warp(_ptr(frames), _ptr(frame_desc), frame_desc.shape[0], _ptr(matrices), _ptr(face_frame), n,
                                     1 if swap_rb else 0, _ptr(u8) if u8 is not None else None, _ptr(f16) if f16 is not None else None,
                                     _lib.stream_ptr()))
    if output == "uint8":
        return u8.cpu().numpy()
    return f16 if output == "device" else (u8, f16)


def frames_to_device(frames_np):
    """[F,H,W,3] uint8 numpy (or list of HxWx3 arrays of differing sizes) -> (cuda uint8 flat, cuda int64 desc [F,4])."""
    torch = _torch()
    if isinstance(frames_np, np.ndarray) and frames_np.ndim == 4:
        frames_np = list(frames_np)
    descs, chunks, off = [], [], 0
    for f in frames_np:
        f = np.ascontiguousarray(f, dtype=np.uint8)
        assert f.ndim == 3 and f.shape[2] == 3
        descs.append((off, f.shape[0], f.shape[1], f.shape[1] * 3))
        chunks.append(f.reshape(-1))
        off += f.size
    flat = torch.from_numpy(np.concatenate(chunks)).cuda()
    desc = torch.tensor(descs, dtype=torch.int64).cuda()
    return flat, desc
