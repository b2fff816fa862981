"""Drop-in for the reference's top-level ``facenet_gpu.py`` (facenet_gpu.py:1-145).

Same names, arguments and error behaviour; the onnxruntime session is replaced by the sm_100a
engine (fire_b200.engine.FaceNetEngine).  `mode` strings are validated exactly like the
reference (facenet_gpu.py:43-60) and then ignored: there is one execution path, the B200 one.

Weights: the reference resolves ``weights/facenet{128,512}.onnx`` relative to its own module
directory (facenet_gpu.py:30-33).  Here the directory is $FIRE_B200_WEIGHTS_DIR (default: the
current working directory, which is the reference checkout when FIRE runs ``python main.py``).
The reference checkout ships git-LFS pointers instead of the real files; a missing file raises
FileNotFoundError and a pointer raises ValueError like a corrupt model would (facenet_gpu.py:36-37,
75-79), unless FIRE_B200_SYNTHETIC_WEIGHTS=1 selects the seeded synthetic tensors.
"""
from __future__ import annotations

import os

import numpy as np

from . import engine as _engine

_VALID_MODES = ("gpu", "gpu_optimized", "cpu", "cpu_optimized", "npu", "npu_optimized")
_PROVIDER = "FireB200Sm100aExecutionProvider"


def check_available_providers():
    """facenet_gpu.py:6-12: print and return the available execution providers."""
    available_providers = [_PROVIDER]
    print("\nAvailable Execution Providers:", available_providers)
    return available_providers


class _IOName:
    def __init__(self, name):
        self.name = name


class FireSession:
    """The subset of onnxruntime.InferenceSession the reference touches (facenet_gpu.py:113-114,127)."""

    def __init__(self, D: int, tensors, source: str):
        self.engine = _engine.FaceNetEngine(D, tensors)
        self.source = source
        self._in, self._out = "input_1", "Bottleneck_BatchNorm"

    def get_inputs(self):
        return [_IOName(self._in)]

    def get_outputs(self):
        return [_IOName(self._out)]

    def get_providers(self):
        return [_PROVIDER]

    def run(self, output_names, feed):
        import torch
        (img,) = feed.values()
        img = np.ascontiguousarray(img, dtype=np.float32)
        if img.ndim != 4 or img.shape[1:] != (160, 160, 3):
            raise ValueError(f"Got invalid dimensions for input: {self._in} expected [N,160,160,3], got {list(img.shape)}")
        x = torch.from_numpy(img).cuda(non_blocking=False)
        raw, _ = self.engine.encode_unit_f32(x, want_l2=False)
        return [raw.cpu().numpy()]


def _weights_dir() -> str:
    return os.environ.get("FIRE_B200_WEIGHTS_DIR", os.getcwd())


def load_facenet_model(onnx_model_path="weights/facenet128.onnx", mode="gpu_optimized"):
    """facenet_gpu.py:14-81 with the session replaced by the sm_100a engine."""
    available_providers = check_available_providers()
    full_onnx_path = os.path.join(_weights_dir(), onnx_model_path)
    D = 512 if "512" in os.path.basename(onnx_model_path) else 128
    synthetic = os.environ.get("FIRE_B200_SYNTHETIC_WEIGHTS", "0") == "1"

    if mode not in _VALID_MODES:
        raise ValueError(f"Invalid mode selected: {mode}. Choose from 'gpu', 'gpu_optimized', 'cpu', 'cpu_optimized', "
                         "'npu', 'npu_optimized'.")
    if not available_providers:
        raise ValueError("None of the desired providers are available on this system.")

    if synthetic:
        from . import weights
        print(f"\nFIRE_B200_SYNTHETIC_WEIGHTS=1: using seeded synthetic FaceNet-{D} weights.\n")
        return FireSession(D, weights.synthetic_weights(D), "synthetic")

    if not os.path.exists(full_onnx_path):
        raise FileNotFoundError(f"ONNX model not found at {full_onnx_path}. Please ensure the file exists.")
    try:
        print(f"\nLoading ONNX model from {full_onnx_path}...\n")
        from . import onnx_reader
        tensors = onnx_reader.load_facenet_tensors(full_onnx_path, D)
        session = FireSession(D, tensors, full_onnx_path)
        print("\nONNX model successfully loaded.\n")
        print("Using Execution Providers:", session.get_providers())
    except Exception as err:
        raise ValueError(f"An error occurred while loading the ONNX model from {full_onnx_path}. "
                         "Please ensure the file is correct and not corrupted.") from err
    return session


class FaceNetClient:
    """facenet_gpu.py:84-129."""

    def __init__(self, model_type="128", mode="gpu"):
        if model_type == "512":
            onnx_model_path = "weights/facenet512.onnx"
            self.output_shape = 512
            self.model_name = "FaceNet-512d"
        else:
            onnx_model_path = "weights/facenet128.onnx"
            self.output_shape = 128
            self.model_name = "FaceNet-128d"
        self.model = load_facenet_model(onnx_model_path=onnx_model_path, mode=mode)
        self.input_shape = (160, 160)
        self.input_name = self.model.get_inputs()[0].name
        self.output_name = self.model.get_outputs()[0].name

    def __call__(self, img: np.ndarray) -> np.ndarray:
        return self.model.run([self.output_name], {self.input_name: img})[0]

    @property
    def engine(self):
        return self.model.engine


def scaling(x, scale):
    """facenet_gpu.py:132-143: the Keras Lambda residual scale (folded into the up-conv weights here)."""
    return x * scale
