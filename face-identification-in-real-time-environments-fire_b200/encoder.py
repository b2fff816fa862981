"""Drop-in for the reference's ``modules/encoder.py`` (modules/encoder.py:9-27).

`Encoder(encoder_model_type, encoder_mode)` keeps `.encoder`, `.input_shape == (160, 160)`,
`.output_shape`, `.encode(face_img)` and `.preprocess_for_encoder(face_img)` with the reference's
argument meaning and error behaviour.  Both run on the B200:

  preprocess_for_encoder  -> fire_preprocess (FIRE_PRE_REFERENCE: cv2 INTER_AREA bit-exact, /255)
  encode                  -> fire_ingest_f32 + fire_facenet_forward

Additive, batched entry points (what bench.py and a batching caller use):
  encode_crops(frames, boxes, box_frame)   frames+boxes in, embeddings out, one launch chain.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _lib, engine as _engine
from .facenet_gpu import FaceNetClient


class Encoder:
    def __init__(self, encoder_model_type: str, encoder_mode: str):
        self.encoder = FaceNetClient(model_type=encoder_model_type, mode=encoder_mode)
        self.input_shape = self.encoder.input_shape
        self.output_shape = self.encoder.output_shape
        logging.info(f"Initialized FaceNet-{self.encoder.output_shape} encoder in {encoder_mode} mode.")

    # ---- reference surface -----------------------------------------------------------------
    def encode(self, face_img: np.ndarray) -> np.ndarray:
        """[B,160,160,3] float32 in [0,1] (what preprocess_for_encoder returns) -> [B,D] float32, un-normalised."""
        return self.encoder(face_img)

    def preprocess_for_encoder(self, face_img: np.ndarray) -> np.ndarray:
        """uint8 [h,w,3] crop -> float32 [1,160,160,3] in [0,1]; ValueError on anything that is not HxWx3."""
        face_img = np.asarray(face_img)
        if face_img.ndim != 3 or face_img.shape[2] != 3 or face_img.shape[0] == 0 or face_img.shape[1] == 0:
            raise ValueError("Face image has incorrect shape for encoder.")
        if face_img.dtype != np.uint8:
            raise ValueError("Face image has incorrect shape for encoder.")   # reference path is uint8-only (cv2 frames)
        import torch
        h, w = face_img.shape[:2]
        flat, desc = _engine.frames_to_device([face_img])
        boxes = torch.tensor([[0, 0, w, h]], dtype=torch.int32, device=flat.device)
        bf = torch.zeros(1, dtype=torch.int32, device=flat.device)
        _, f32, _ = _engine.preprocess_boxes(flat, desc, boxes, bf, _lib.PRE_REFERENCE, want_f16=False, want_f32=True)
        return f32.cpu().numpy()

    # ---- batched additions -------------------------------------------------------------------
    def encode_crops(self, frames, boxes, box_frame=None, mode: int = _lib.PRE_REFERENCE, normalize: bool = False):
        """frames: uint8 [F,H,W,3] (numpy or list of arrays); boxes: int [n,4] xywh as the tracker emits them;
        box_frame: int [n] (default all zeros).  Returns (embeddings [n,D] float32 numpy, status [n] int32)
        where status 1 marks an empty crop (the reference skips those faces, face_recognition.py:418-420)."""
        import torch
        flat, desc = _engine.frames_to_device(frames)
        boxes_t = torch.as_tensor(np.asarray(boxes, dtype=np.int32).reshape(-1, 4)).to(flat.device)
        n = boxes_t.shape[0]
        bf = torch.zeros(n, dtype=torch.int32, device=flat.device) if box_frame is None else \
            torch.as_tensor(np.asarray(box_frame, dtype=np.int32)).to(flat.device)
        f16, _, status = _engine.preprocess_boxes(flat, desc, boxes_t, bf, mode, want_f16=True, want_f32=False)
        raw, l2 = self.encoder.engine.forward(f16, want_l2=normalize)
        out = (l2 if normalize else raw).cpu().numpy()
        return out, status.cpu().numpy()
