"""Minimal ONNX wire-format reader for weights/facenet{128,512}.onnx (no `onnx` package needed).

The reference loads these files with onnxruntime (facenet_gpu.py:72).  The engine only needs the
tensors, so this walks the protobuf by hand: ModelProto.graph(7) -> GraphProto.node(1) /
initializer(5); NodeProto input(1) output(2) name(3) op_type(4) attribute(5); TensorProto dims(1)
data_type(2) float_data(4) name(8) raw_data(9).

`load_facenet_tensors` maps the graph onto the Keras tensor names the plan consumes
(netplan.Plan.keras_tensor_shapes).  Matching is by layer name (tf2onnx / keras2onnx keep the Keras
layer name inside node and initializer names), then verified by shape.  Exports that constant-fold
BatchNorm into the convolution are accepted: the folded bias is expressed as an identity BN.

STATUS: the real files are git-LFS pointers in the reference checkout, so this path is
exercised only against ONNX files written by tests/onnx_writer.py.
"""
from __future__ import annotations

import re
import struct
from typing import Dict, List, Tuple

import numpy as np

from .netplan import BN_EPS, Plan


# ---- protobuf wire format ----------------------------------------------------------------------------
def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: memoryview):
    """Yield (field_number, wire_type, value) - value is int for varint/fixed, memoryview for len-delimited."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _packed_varints(v) -> List[int]:
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims, dtype, name, raw, floats = [], 1, "", None, []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            dims.extend(_packed_varints(v) if wt == 2 else [v])
        elif fn == 2:
            dtype = v
        elif fn == 4:
            floats.append(np.frombuffer(v, dtype="<f4") if wt == 2 else np.array([struct.unpack("<f", struct.pack("<I", v))[0]], "<f4"))
        elif fn == 8:
            name = bytes(v).decode()
        elif fn == 9:
            raw = v
    np_dtype = {1: "<f4", 6: "<i4", 7: "<i8", 10: "<f2", 11: "<f8"}.get(dtype)
    if np_dtype is None:
        return name, np.zeros(0, np.float32)
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dtype)
    elif floats:
        arr = np.concatenate(floats)
    else:
        arr = np.zeros(0, np_dtype)
    shape = [int(d) for d in dims]
    if int(np.prod(shape)) == arr.size:
        arr = arr.reshape(shape)
    return name, arr


def parse_model(path: str):
    """-> (nodes, initializers): nodes = list of dicts {op, name, inputs, outputs}; initializers = {name: ndarray}."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    if len(data) < 1024 and bytes(data[:7]) == b"version":
        raise ValueError(f"{path} is a git-LFS pointer, not an ONNX model")
    graph = None
    for fn, wt, v in _fields(data):
        if fn == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found")
    nodes, inits = [], {}
    for fn, wt, v in _fields(graph):
        if fn == 1 and wt == 2:
            node = {"op": "", "name": "", "inputs": [], "outputs": []}
            for f2, w2, x in _fields(v):
                if f2 == 1:
                    node["inputs"].append(bytes(x).decode())
                elif f2 == 2:
                    node["outputs"].append(bytes(x).decode())
                elif f2 == 3:
                    node["name"] = bytes(x).decode()
                elif f2 == 4:
                    node["op"] = bytes(x).decode()
            nodes.append(node)
        elif fn == 5 and wt == 2:
            name, arr = _tensor(v)
            inits[name] = arr
    return nodes, inits


# ---- mapping onto the plan's tensor names ----------------------------------------------------------------
def _has_component(haystack: str, layer: str) -> bool:
    return re.search(r"(^|[/:_.])" + re.escape(layer) + r"($|[/:.])", haystack) is not None


def load_facenet_tensors(path: str, D: int) -> Dict[str, np.ndarray]:
    nodes, inits = parse_model(path)
    want = Plan(D, fuse_siblings=False).keras_tensor_shapes()
    layers = sorted({k.split("/")[0] for k in want if k.endswith("/kernel")}, key=len, reverse=True)
    out: Dict[str, np.ndarray] = {}
    convs = [n for n in nodes if n["op"] in ("Conv", "MatMul", "Gemm")]
    bns = [n for n in nodes if n["op"] == "BatchNormalization"]

    def find(cands, layer):
        hits = [n for n in cands if _has_component(n["name"], layer) or any(_has_component(i, layer) for i in n["inputs"])]
        return hits[0] if hits else None

    for layer in layers:
        shape = want[layer + "/kernel"]
        node = find(convs, layer)
        if node is None:
            raise ValueError(f"ONNX graph has no Conv/MatMul for Keras layer {layer!r}")
        consts = [inits[i] for i in node["inputs"] if i in inits]
        w = next((c for c in consts if c.ndim in (2, 4)), None)
        if w is None:
            raise ValueError(f"no weight initializer for {layer!r}")
        w = np.asarray(w, dtype=np.float32)
        if w.ndim == 4:                                  # ONNX Conv weights are OIHW -> Keras HWIO
            w = w.transpose(2, 3, 1, 0)
        elif node["op"] == "Gemm" and w.shape != tuple(shape):
            w = w.T
        if tuple(w.shape) != tuple(shape):
            raise ValueError(f"{layer}: weight shape {w.shape} != expected {shape}")
        out[layer + "/kernel"] = np.ascontiguousarray(w)
        bias = next((np.asarray(c, np.float32) for c in consts if c.ndim == 1 and c.shape[0] == shape[-1]), None)
        if layer + "/bias" in want:
            if bias is None:
                raise ValueError(f"{layer}: expected a bias")
            out[layer + "/bias"] = bias
            continue
        bn = find(bns, layer + "_BatchNorm")
        if bn is not None:
            scale, beta, mean, var = (np.asarray(inits[i], np.float32) for i in bn["inputs"][1:5])
            # scale=False in the Keras model -> gamma == 1; a non-unit gamma is folded into mean/var/beta exactly
            inv = scale / np.sqrt(var.astype(np.float64) + BN_EPS)
            out[layer + "_BatchNorm/beta"] = beta
            out[layer + "_BatchNorm/moving_mean"] = mean
            out[layer + "_BatchNorm/moving_variance"] = (1.0 / (inv * inv) - BN_EPS).astype(np.float32) \
                if not np.allclose(scale, 1.0) else var
            if not np.allclose(scale, 1.0) and np.any(scale < 0):
                raise ValueError(f"{layer}: negative BatchNorm gamma is not representable with scale=False")
        else:                                            # BN constant-folded into the conv: identity BN + folded bias
            n = shape[-1]
            out[layer + "_BatchNorm/beta"] = bias if bias is not None else np.zeros(n, np.float32)
            out[layer + "_BatchNorm/moving_mean"] = np.zeros(n, np.float32)
            out[layer + "_BatchNorm/moving_variance"] = np.full(n, 1.0 - BN_EPS, np.float32)
    missing = [k for k in want if k not in out]
    if missing:
        raise ValueError(f"ONNX model lacks tensors for: {missing[:5]} ...")
    return out
