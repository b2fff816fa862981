"""Drop-in for the reference's ``processing/preprocess.py`` (processing/preprocess.py:10-145) plus the
batched GPU entry point the north star adds.

The five reference functions are detector-side loaders (RetinaFace only); they are not on
the hot path and stay host-side, with the same names, arguments, return values and errors:
    get_image, load_base64_img, load_image_from_web, resize_image, preprocess_image

New (GPU, fire_preprocess in include/fire_b200.h):
    crop_resize_normalize(frames, boxes, box_frame=None, mode="reference"|"northstar", ...)
"""
from __future__ import annotations

import base64
import os
from pathlib import Path
from typing import Union

import cv2
import numpy as np

from . import _lib


# ---- reference API (host side) ---------------------------------------------------------------------
def get_image(img_uri: Union[str, np.ndarray]) -> np.ndarray:
    """Accepts a BGR numpy array, a ``data:image/`` base64 string, an http(s) URL or a file path."""
    if isinstance(img_uri, np.ndarray):
        img = img_uri.copy()
    elif isinstance(img_uri, str) and img_uri.startswith("data:image/"):
        img = load_base64_img(img_uri)
    elif isinstance(img_uri, str) and img_uri.startswith("http"):
        img = load_image_from_web(url=img_uri)
    elif isinstance(img_uri, (str, Path)):
        path = str(img_uri)
        if not os.path.isfile(path):
            raise ValueError(f"Input image file path ({path}) does not exist.")
        img = cv2.imread(path)
    else:
        raise ValueError(f"Invalid image input - {img_uri}."
                         "Exact paths, pre-loaded numpy arrays, base64 encoded strings and urls are welcome.")
    if len(img.shape) != 3 or np.prod(img.shape) == 0:
        raise ValueError("Input image needs to have 3 channels at must not be empty.")
    return img


def load_base64_img(uri) -> np.ndarray:
    payload = uri.split(",")[1]
    raw = np.frombuffer(base64.b64decode(payload), dtype=np.uint8)    # np.fromstring is gone in numpy 2
    return cv2.imdecode(raw, cv2.IMREAD_COLOR)


def load_image_from_web(url: str) -> np.ndarray:
    import requests
    response = requests.get(url, stream=True, timeout=60)
    response.raise_for_status()
    data = np.asarray(bytearray(response.raw.read()), dtype=np.uint8)
    return cv2.imdecode(data, cv2.IMREAD_COLOR)


def resize_image(img: np.ndarray, scales: list, allow_upscaling: bool) -> tuple:
    """Scale so the short side hits scales[0] without the long side exceeding scales[1] (INTER_LINEAR)."""
    h, w = img.shape[0:2]
    short, long_ = (h, w) if w > h else (w, h)
    im_scale = scales[0] / float(short)
    if not allow_upscaling:
        im_scale = min(1.0, im_scale)
    if np.round(im_scale * long_) > scales[1]:
        im_scale = scales[1] / float(long_)
    if im_scale != 1.0:
        img = cv2.resize(img, None, None, fx=im_scale, fy=im_scale, interpolation=cv2.INTER_LINEAR)
    return img, im_scale


def preprocess_image(img: np.ndarray, allow_upscaling: bool) -> tuple:
    """-> (float32 [1,H,W,3] RGB tensor, (H, W), im_scale); means 0, stds 1, scales [1024, 1980]."""
    img, im_scale = resize_image(img, [1024, 1980], allow_upscaling)
    rgb = img.astype(np.float32)[:, :, ::-1]
    im_tensor = np.ascontiguousarray(rgb)[None, ...]
    return im_tensor, img.shape[0:2], im_scale


# ---- GPU batch preprocessing ---------------------------------------------------------------------
_MODES = {"reference": _lib.PRE_REFERENCE, "northstar": _lib.PRE_NORTHSTAR}


def crop_resize_normalize(frames, boxes, box_frame=None, mode: str = "reference", swap_rb: bool = False,
                          output: str = "float32"):
    """Crop every box out of its frame (reference clamp rule, face_recognition.py:412-420), resize to 160x160
    and normalise, for the whole batch in one kernel launch on the B200.

    frames : uint8 [F,H,W,3] array or list of HxWx3 arrays (BGR or RGB - channels are not interpreted)
    boxes  : int [n,4] x, y, w, h
    mode   : "reference" = cv2.resize(INTER_AREA) on uint8 then /255 (bit-exact with modules/encoder.py:19-27)
             "northstar" = float bilinear + per-crop prewhiten
    output : "float32" -> numpy [n,160,160,3] (what preprocess_for_encoder returns, stacked)
             "device"  -> (cuda fp16 [n,160,160,8] network input, cuda int32 status[n]) for Encoder/engine use
    Returns (array, status) where status[i] == 1 marks an empty crop (all-zero output).
    """
    import torch
    from . import engine
    if mode not in _MODES:
        raise ValueError(f"mode must be one of {sorted(_MODES)}, not {mode!r}")
    m = _MODES[mode] | (_lib.PRE_FLAG_SWAP_RB if swap_rb else 0)
    flat, desc = engine.frames_to_device(frames)
    boxes_t = torch.as_tensor(np.asarray(boxes, dtype=np.int32).reshape(-1, 4)).to(flat.device)
    n = boxes_t.shape[0]
    bf = torch.zeros(n, dtype=torch.int32, device=flat.device) if box_frame is None else \
        torch.as_tensor(np.asarray(box_frame, dtype=np.int32)).to(flat.device)
    if output == "device":
        f16, _, status = engine.preprocess_boxes(flat, desc, boxes_t, bf, m, want_f16=True, want_f32=False)
        return f16, status
    _, f32, status = engine.preprocess_boxes(flat, desc, boxes_t, bf, m, want_f16=False, want_f32=True)
    return f32.cpu().numpy(), status.cpu().numpy()


# Alignment targets of the reference's extract_faces: left eye, right eye, nose tip in the 160 x 160 crop
# (yunet_face_detector.py:143-147; identical in the MediaPipe and RetinaFace detectors)
ALIGN_DST = np.float32([(0.35 * 160, 0.35 * 160), (0.65 * 160, 0.35 * 160), (0.5 * 160, 0.55 * 160)])


def get_affine_transform(src_pts, dst_pts) -> np.ndarray:
    """cv2.getAffineTransform bit for bit: the 6 x 6 system solved with OpenCV's LU elimination (partial pivoting,
    `d = -1/a_ii; row_j += (a_ji * d) * row_i`, back substitution `s -= a_ik * x_k; x_i = s / a_ii`), all in float64 on
    float32 inputs.  Returns the forward 2 x 3 matrix (float64)."""
    s = np.asarray(src_pts, dtype=np.float32).reshape(3, 2)
    d = np.asarray(dst_pts, dtype=np.float32).reshape(3, 2)
    a = [[0.0] * 6 for _ in range(6)]
    b = [0.0] * 6
    for i in range(3):
        x, y = float(s[i, 0]), float(s[i, 1])
        a[2 * i][0:3] = [x, y, 1.0]
        a[2 * i + 1][3:6] = [x, y, 1.0]
        b[2 * i], b[2 * i + 1] = float(d[i, 0]), float(d[i, 1])
    m, eps = 6, 2.220446049250313e-16 * 100
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(a[j][i]) > abs(a[k][i]):
                k = j
        if abs(a[k][i]) < eps:
            return np.zeros((2, 3), dtype=np.float64)
        if k != i:
            a[i][i:], a[k][i:] = a[k][i:], a[i][i:]
            b[i], b[k] = b[k], b[i]
        dd = -1 / a[i][i]
        for j in range(i + 1, m):
            alpha = a[j][i] * dd
            for kk in range(i + 1, m):
                a[j][kk] += alpha * a[i][kk]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        acc = b[i]
        for k in range(i + 1, m):
            acc -= a[i][k] * b[k]
        b[i] = acc / a[i][i]
    return np.array(b, dtype=np.float64).reshape(2, 3)


def align_faces(frames, landmarks, face_frame=None, swap_rb: bool = True, output: str = "uint8"):
    """Aligned 160 x 160 crops of the enrol path, on the B200: for every face
        M = getAffineTransform([left_eye, right_eye, nose], ALIGN_DST);  crop = warpAffine(frame, M, (160, 160))[:, :, ::-1]
    exactly as the reference's extract_faces(align=True) does with cv2 (yunet_face_detector.py:135-165).

    frames    : uint8 [F,H,W,3] array or list of HxWx3 arrays
    landmarks : [n,3,2] points in the order the reference passes them to getAffineTransform (left eye, right eye, nose)
    swap_rb   : apply the final [:, :, ::-1] (True = what extract_faces returns)
    output    : "uint8" -> numpy uint8 [n,160,160,3];  "device" -> cuda fp16 network input [n,80,80,16] for the engine
    """
    import torch
    from . import engine
    lm = np.asarray(landmarks, dtype=np.float32).reshape(-1, 3, 2)
    n = lm.shape[0]
    mats = np.stack([get_affine_transform(lm[i], ALIGN_DST) for i in range(n)]).reshape(n, 6)
    flat, desc = engine.frames_to_device(frames)
    ff = np.zeros(n, dtype=np.int32) if face_frame is None else np.asarray(face_frame, dtype=np.int32)
    return engine.align_warp(flat, desc, torch.from_numpy(mats).to(flat.device), torch.from_numpy(ff).to(flat.device), swap_rb, output)
