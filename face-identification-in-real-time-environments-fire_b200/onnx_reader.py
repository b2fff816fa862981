"""Minimal ONNX wire-format reader for weights/facenet{128,512}.onnx (no `onnx` package needed).

The reference loads these files with onnxruntime (facenet_gpu.py:72).  The engine only needs the
tensors, so this walks the protobuf by hand: ModelProto.graph(7) -> GraphProto.node(1) /
initializer(5); NodeProto input(1) output(2) name(3) op_type(4) attribute(5); TensorProto dims(1)
data_type(2) float_data(4) name(8) raw_data(9).

`load_facenet_tensors` maps the graph onto the Keras tensor names the plan consumes
(netplan.Plan.keras_tensor_shapes).  Matching is by layer name (tf2onnx / keras2onnx keep the Keras
layer name inside node and initializer names), then verified by shape.  Exports that constant-fold
BatchNorm into the convolution are accepted: the folded bias is expressed as an identity BN.

The graph is also checked structurally (`verify_facenet_graph`): operator set, 132 Conv + 1 Dense, 23 Concat, 3 MaxPool and
the 21 residual scales of the `scaling` Lambda (facenet_gpu.py:132-143) read from the Mul constants.  Whatever per-channel
element-wise ops an exporter leaves behind a convolution (BatchNormalization, BN folded into the conv, BN decomposed into
Mul / Add / Sub) are composed into one affine map and re-expressed as the scale=False BatchNorm the plan folds.

STATUS: the real files are git-LFS pointers in the reference checkout.  The path is exercised against files
with the full branch structure written by tests/onnx_graph_writer.py - the same files cv2.dnn.readNetFromONNX executes as
the independent check of the FaceNet oracle - and against the linear chain of tests/onnx_writer.py.
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


ALLOWED_OPS = {"Transpose", "Conv", "BatchNormalization", "Relu", "MaxPool", "Concat", "Mul", "Add", "Sub", "GlobalAveragePool", "ReduceMean",
               "MatMul", "Gemm", "Reshape", "Squeeze", "Flatten", "Identity", "Dropout", "Unsqueeze", "Cast", "Shape", "Gather", "Constant"}
RESIDUAL_SCALES = {"Block35": 0.17, "Block17": 0.1, "Block8": 0.2}      # facenet_gpu.py:132-143 `scaling`; Block8_6 uses 1.0


def _per_channel(c: np.ndarray, n: int):
    """A constant operand that is one value per output channel ([n], [1,n], [n,1,1], [1,n,1,1]) -> flat [n], else None."""
    c = np.asarray(c)
    if c.size != n:
        return None
    if c.ndim <= 1 or sorted(c.shape)[-1] == n:
        return c.reshape(n).astype(np.float64)
    return None


def _affine_chain(node, consumers, inits, n: int):
    """Follow the output of a Conv/MatMul through the element-wise ops with a constant per-channel operand that exporters put
    behind it (BatchNormalization, or its decomposed Mul/Add/Sub form).  Returns (a, b, ops, last_node, scalar_mul): the
    chain computes a * x + b per channel; `scalar_mul` is the constant of a following Mul by ONE scalar (the residual
    `scaling` Lambda), which is not part of the layer."""
    a, b, ops = np.ones(n), np.zeros(n), []
    cur, scalar_mul = node, None
    while True:
        nxt = consumers.get(cur["outputs"][0], [])
        if len(nxt) != 1:
            break
        m = nxt[0]
        consts = [(i, inits[i]) for i in m["inputs"] if i in inits]
        if m["op"] == "BatchNormalization" and len(consts) == 4:
            g, beta, mean, var = (np.asarray(inits[i], np.float64) for i in m["inputs"][1:5])
            inv = g / np.sqrt(var + BN_EPS)
            a, b = a * inv, (b - mean) * inv + beta
        elif m["op"] in ("Mul", "Add", "Sub") and len(consts) == 1:
            c = np.asarray(consts[0][1])
            if m["op"] == "Mul" and c.size == 1 and n > 1:
                scalar_mul = float(c.reshape(()))
                break
            v = _per_channel(c, n)
            if v is None:
                break
            if m["op"] == "Mul":
                a, b = a * v, b * v
            elif m["op"] == "Add":
                b = b + v
            elif m["inputs"][0] == cur["outputs"][0]:        # x - c
                b = b - v
            else:                                            # c - x
                a, b = -a, v - b
        else:
            break
        ops.append(m)
        cur = m
    return a, b, ops, cur, scalar_mul


def verify_facenet_graph(nodes, D: int, scales: Dict[str, float]) -> None:
    """The structure SURVEY App. A describes, checked on the parsed graph: op set, 132 convolutions + one Dense, 23 concats,
    3 max-pools, and the 21 residual scales (0.17 x 5, 0.1 x 10, 0.2 x 5, 1.0 x 1)."""
    from collections import Counter
    ops = Counter(n["op"] for n in nodes)
    unknown = set(ops) - ALLOWED_OPS
    if unknown:
        raise ValueError(f"ONNX graph uses operators outside the FaceNet set: {sorted(unknown)}")
    if ops["Conv"] != 132 or ops["MatMul"] + ops["Gemm"] != 1:
        raise ValueError(f"ONNX graph has {ops['Conv']} Conv and {ops['MatMul'] + ops['Gemm']} Dense nodes, expected 132 and 1")
    if ops["Concat"] != 23 or ops["MaxPool"] != 3:
        raise ValueError(f"ONNX graph has {ops['Concat']} Concat / {ops['MaxPool']} MaxPool nodes, expected 23 / 3")
    want = {f"{fam}_{i}_Conv2d_1x1": (1.0 if (fam == "Block8" and i == 6) else sc)
            for fam, cnt, sc in (("Block35", 5, 0.17), ("Block17", 10, 0.1), ("Block8", 6, 0.2)) for i in range(1, cnt + 1)}
    for layer, sc in want.items():
        got = scales.get(layer, 1.0)           # an exporter may drop a Mul by 1.0
        if abs(got - sc) > 1e-6:
            raise ValueError(f"{layer}: residual scale {got} in the ONNX graph, {sc} expected (facenet_gpu.py:132-143)")


def load_facenet_tensors(path: str, D: int, verify: bool = True) -> Dict[str, np.ndarray]:
    nodes, inits = parse_model(path)
    want = Plan(D, fuse_siblings=False).keras_tensor_shapes()
    layers = sorted({k.split("/")[0] for k in want if k.endswith("/kernel")}, key=len, reverse=True)
    out: Dict[str, np.ndarray] = {}
    convs = [n for n in nodes if n["op"] in ("Conv", "MatMul", "Gemm")]
    consumers: Dict[str, list] = {}
    for n in nodes:
        for i in n["inputs"]:
            consumers.setdefault(i, []).append(n)
    scales: Dict[str, float] = {}

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
        n = shape[-1]
        bias = next((np.asarray(c, np.float32) for c in consts if c.ndim == 1 and c.shape[0] == n), None)
        a, b, chain, _, scalar_mul = _affine_chain(node, consumers, inits, n)
        if layer + "/bias" in want:                      # residual `up` conv: bias, no BN; the scalar Mul behind it is the block's scale
            if bias is None:
                raise ValueError(f"{layer}: expected a bias")
            if chain:
                raise ValueError(f"{layer}: unexpected per-channel ops behind a residual up-convolution")
            out[layer + "/bias"] = bias
            if scalar_mul is not None:
                scales[layer] = scalar_mul
            continue
        if len(chain) == 1 and chain[0]["op"] == "BatchNormalization" and bias is None and \
                np.allclose(np.asarray(inits[chain[0]["inputs"][1]]), 1.0):
            _, beta, mean, var = (np.asarray(inits[i], np.float32) for i in chain[0]["inputs"][1:5])      # the Keras tensors, verbatim
            out[layer + "_BatchNorm/beta"], out[layer + "_BatchNorm/moving_mean"], out[layer + "_BatchNorm/moving_variance"] = beta, mean, var
            continue
        # anything else (BN folded into the conv, gamma != 1, BN decomposed into Mul/Add/Sub): y = a * (conv + bias) + b per
        # channel, re-expressed as the scale=False BatchNorm the plan folds: 1/sqrt(var + eps) = a, beta = a * bias + b, mean = 0
        if np.any(a <= 0):
            raise ValueError(f"{layer}: non-positive BatchNorm scale is not representable with scale=False")
        b0 = np.zeros(n) if bias is None else bias.astype(np.float64)
        out[layer + "_BatchNorm/beta"] = (a * b0 + b).astype(np.float32)
        out[layer + "_BatchNorm/moving_mean"] = np.zeros(n, np.float32)
        out[layer + "_BatchNorm/moving_variance"] = (1.0 / (a * a) - BN_EPS).astype(np.float32)
    missing = [k for k in want if k not in out]
    if missing:
        raise ValueError(f"ONNX model lacks tensors for: {missing[:5]} ...")
    if verify:
        verify_facenet_graph(nodes, D, scales)
    return out
