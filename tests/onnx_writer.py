"""Minimal ONNX (protobuf wire format) writer used to exercise fire_b200.onnx_reader without the `onnx` package:
writes a Conv -> BatchNormalization -> Relu ... MatMul graph whose node/initializer names carry the Keras layer
names the way tf2onnx exports do."""
import struct

import numpy as np


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _vi(field, value):
    return _varint(field << 3) + _varint(value)


def tensor(name, arr):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    body = b"".join(_vi(1, d) for d in arr.shape) + _vi(2, 1) + _ld(8, name.encode()) + _ld(9, arr.tobytes())
    return body


def node(op, name, inputs, outputs):
    return b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs) + \
        _ld(3, name.encode()) + _ld(4, op.encode())


def write_facenet_like(path, tensors, fold_bn=False):
    """tensors: Keras-named dict (fire_b200.weights.synthetic_weights).  fold_bn=True emits Conv+bias only."""
    nodes, inits = [], []
    layers = [k[:-len("/kernel")] for k in tensors if k.endswith("/kernel")]
    prev = "input_1"
    for layer in layers:
        k = tensors[layer + "/kernel"]
        out = f"model/{layer}/out:0"
        if k.ndim == 2:
            inits.append(tensor(f"model/{layer}/MatMul/ReadVariableOp:0", k))
            nodes.append(node("MatMul", f"model/{layer}/MatMul", [prev, f"model/{layer}/MatMul/ReadVariableOp:0"], [out]))
        else:
            w = k.transpose(3, 2, 0, 1)                                  # HWIO -> OIHW
            ins = [prev, f"model/{layer}/Conv2D/ReadVariableOp:0"]
            bias = tensors.get(layer + "/bias")
            if bias is None and fold_bn:
                inv = 1.0 / np.sqrt(tensors[layer + "_BatchNorm/moving_variance"].astype(np.float64) + 1e-3)
                w = (w * inv[:, None, None, None]).astype(np.float32)
                bias = (tensors[layer + "_BatchNorm/beta"] - tensors[layer + "_BatchNorm/moving_mean"] * inv).astype(np.float32)
            inits.append(tensor(ins[1], w))
            if bias is not None:
                inits.append(tensor(f"model/{layer}/BiasAdd/ReadVariableOp:0", bias))
                ins.append(f"model/{layer}/BiasAdd/ReadVariableOp:0")
            nodes.append(node("Conv", f"model/{layer}/Conv2D", ins, [out]))
        prev = out
        if layer + "_BatchNorm/beta" in tensors and not (fold_bn and k.ndim == 4):
            bn = layer + "_BatchNorm"
            names = [f"model/{bn}/{p}:0" for p in ("gamma", "beta", "mean", "var")]
            vals = [np.ones_like(tensors[bn + "/beta"]), tensors[bn + "/beta"], tensors[bn + "/moving_mean"], tensors[bn + "/moving_variance"]]
            for n_, v in zip(names, vals):
                inits.append(tensor(n_, v))
            out = f"model/{bn}/out:0"
            nodes.append(node("BatchNormalization", f"model/{bn}/FusedBatchNormV3", [prev] + names, [out]))
            prev = out
    graph = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"facenet") + b"".join(_ld(5, t) for t in inits)
    model = _vi(1, 8) + _ld(7, graph)
    with open(path, "wb") as f:
        f.write(model)
