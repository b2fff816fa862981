"""Writes the WHOLE FaceNet graph as an ONNX file, the way an exporter emits the deepface/Keras Inception-ResNet-v1
(SURVEY App. A): NHWC input -> Transpose -> Conv / BatchNormalization / Relu / MaxPool, the three branch families with
Concat, the `scaling` Lambda of facenet_gpu.py:132-143 as a Mul by a constant followed by Add (+ Relu), GlobalAveragePool ->
Flatten -> MatMul -> BatchNormalization.  Protobuf wire format by hand (the `onnx` package is not installed).

Test infrastructure.  Two uses:
  * an INDEPENDENT executor (cv2.dnn.readNetFromONNX, the one ONNX runtime in this image) runs the file, which pins
    oracle/facenet_ref.py and the GPU engine against arithmetic neither of them wrote (tests/test_oracle_facenet.py,
    tests/golden/make_facenet_golden.py);
  * fire_b200.onnx_reader parses a file with the real branch structure (tests/test_onnx_reader.py)."""
import struct

import numpy as np


def _varint(n):
    n &= (1 << 64) - 1
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


def _f32(field, value):
    return _varint((field << 3) | 5) + struct.pack("<f", value)


def tensor(name, arr, dtype=np.float32):
    arr = np.ascontiguousarray(arr, dtype=dtype)
    code = {np.dtype(np.float32): 1, np.dtype(np.int64): 7}[arr.dtype]
    return b"".join(_vi(1, d) for d in arr.shape) + _vi(2, code) + _ld(8, name.encode()) + _ld(9, arr.tobytes())


def attr_ints(name, vals):
    return _ld(1, name.encode()) + b"".join(_vi(8, v) for v in vals) + _vi(20, 7)


def attr_int(name, v):
    return _ld(1, name.encode()) + _vi(3, v) + _vi(20, 2)


def attr_float(name, v):
    return _ld(1, name.encode()) + _f32(2, v) + _vi(20, 1)


def node(op, name, inputs, outputs, attrs=()):
    return b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs) + \
        _ld(3, name.encode()) + _ld(4, op.encode()) + b"".join(_ld(5, a) for a in attrs)


def value_info(name, shape):
    dims = b"".join(_ld(1, _vi(1, d)) for d in shape)
    ttype = _ld(1, _vi(1, 1) + _ld(2, dims))
    return _ld(1, name.encode()) + _ld(2, ttype)


class _G:
    def __init__(self, tensors, prefix, fold_bn):
        self.t, self.p, self.fold = tensors, prefix, fold_bn
        self.nodes, self.inits = [], []
        self.n = 0

    def const(self, name, arr, dtype=np.float32):
        self.inits.append(tensor(name, arr, dtype))
        return name

    def op(self, op, name, inputs, attrs=()):
        out = f"{name}:0"
        self.nodes.append(node(op, name, inputs, [out], attrs))
        return out

    def conv(self, x, layer, stride=1, same=False, relu=True):
        k = self.t[layer + "/kernel"]
        kh, kw = k.shape[:2]
        w = k.transpose(3, 2, 0, 1)                                      # HWIO -> OIHW
        pads = [kh // 2, kw // 2, kh // 2, kw // 2] if same else [0, 0, 0, 0]
        ins = [x]
        bias = self.t.get(layer + "/bias")
        has_bn = layer + "_BatchNorm/beta" in self.t
        if has_bn and self.fold:                                          # exporter constant-folded BN into the conv
            inv = 1.0 / np.sqrt(self.t[layer + "_BatchNorm/moving_variance"].astype(np.float64) + 1e-3)
            w = (w * inv[:, None, None, None]).astype(np.float32)
            bias = (self.t[layer + "_BatchNorm/beta"] - self.t[layer + "_BatchNorm/moving_mean"] * inv).astype(np.float32)
        ins.append(self.const(f"{self.p}/{layer}/Conv2D/ReadVariableOp:0", w))
        if bias is not None:
            ins.append(self.const(f"{self.p}/{layer}/BiasAdd/ReadVariableOp:0", bias))
        y = self.op("Conv", f"{self.p}/{layer}/Conv2D", ins,
                    [attr_ints("dilations", [1, 1]), attr_int("group", 1), attr_ints("kernel_shape", [kh, kw]), attr_ints("pads", pads),
                     attr_ints("strides", [stride, stride])])
        if has_bn and not self.fold:
            bn = layer + "_BatchNorm"
            names = [self.const(f"{self.p}/{bn}/{q}:0", v) for q, v in
                     (("ones", np.ones_like(self.t[bn + "/beta"])), ("ReadVariableOp", self.t[bn + "/beta"]),
                      ("FusedBatchNormV3/ReadVariableOp", self.t[bn + "/moving_mean"]), ("FusedBatchNormV3/ReadVariableOp_1", self.t[bn + "/moving_variance"]))]
            y = self.op("BatchNormalization", f"{self.p}/{bn}/FusedBatchNormV3", [y] + names, [attr_float("epsilon", 1e-3), attr_float("momentum", 0.995)])
        if relu:
            y = self.op("Relu", f"{self.p}/{layer}_Activation/Relu", [y])
        return y

    def maxpool(self, x, name):
        return self.op("MaxPool", f"{self.p}/{name}/MaxPool", [x], [attr_ints("kernel_shape", [3, 3]), attr_ints("pads", [0, 0, 0, 0]), attr_ints("strides", [2, 2])])

    def concat(self, xs, name):
        return self.op("Concat", f"{self.p}/{name}/concat", xs, [attr_int("axis", 1)])

    def residual(self, x, branches, p, scale, relu):
        cat = self.concat(branches, f"{p}_Concatenate")
        up = self.conv(cat, f"{p}_Conv2d_1x1", relu=False)
        s = self.const(f"{self.p}/{p}_ScaleSum/mul/y:0", np.array(scale, np.float32))
        scaled = self.op("Mul", f"{self.p}/{p}_ScaleSum/mul", [up, s])        # the `scaling` Lambda: x * scale
        y = self.op("Add", f"{self.p}/{p}_ScaleSum/add", [x, scaled])
        return self.op("Relu", f"{self.p}/{p}_Activation/Relu", [y]) if relu else y


def write_facenet_graph(path, tensors, D, fold_bn=False, prefix="inception_resnet_v1"):
    g = _G(tensors, prefix, fold_bn)
    x = g.op("Transpose", f"{prefix}/Conv2d_1a_3x3/Conv2D__6", ["input_1"], [attr_ints("perm", [0, 3, 1, 2])])
    x = g.conv(x, "Conv2d_1a_3x3", stride=2)
    x = g.conv(x, "Conv2d_2a_3x3")
    x = g.conv(x, "Conv2d_2b_3x3", same=True)
    x = g.maxpool(x, "MaxPool_3a_3x3")
    x = g.conv(x, "Conv2d_3b_1x1")
    x = g.conv(x, "Conv2d_4a_3x3")
    x = g.conv(x, "Conv2d_4b_3x3", stride=2)
    for i in range(1, 6):
        p = f"Block35_{i}"
        b0 = g.conv(x, f"{p}_Branch_0_Conv2d_1x1", same=True)
        b1 = g.conv(g.conv(x, f"{p}_Branch_1_Conv2d_0a_1x1", same=True), f"{p}_Branch_1_Conv2d_0b_3x3", same=True)
        b2 = g.conv(g.conv(g.conv(x, f"{p}_Branch_2_Conv2d_0a_1x1", same=True), f"{p}_Branch_2_Conv2d_0b_3x3", same=True), f"{p}_Branch_2_Conv2d_0c_3x3", same=True)
        x = g.residual(x, [b0, b1, b2], p, 0.17, True)
    b0 = g.conv(x, "Mixed_6a_Branch_0_Conv2d_1a_3x3", stride=2)
    b1 = g.conv(g.conv(g.conv(x, "Mixed_6a_Branch_1_Conv2d_0a_1x1", same=True), "Mixed_6a_Branch_1_Conv2d_0b_3x3", same=True), "Mixed_6a_Branch_1_Conv2d_1a_3x3", stride=2)
    x = g.concat([b0, b1, g.maxpool(x, "Mixed_6a_Branch_2_MaxPool_1a_3x3")], "Mixed_6a")
    for i in range(1, 11):
        p = f"Block17_{i}"
        b0 = g.conv(x, f"{p}_Branch_0_Conv2d_1x1", same=True)
        b1 = g.conv(g.conv(g.conv(x, f"{p}_Branch_1_Conv2d_0a_1x1", same=True), f"{p}_Branch_1_Conv2d_0b_1x7", same=True), f"{p}_Branch_1_Conv2d_0c_7x1", same=True)
        x = g.residual(x, [b0, b1], p, 0.1, True)
    b0 = g.conv(g.conv(x, "Mixed_7a_Branch_0_Conv2d_0a_1x1", same=True), "Mixed_7a_Branch_0_Conv2d_1a_3x3", stride=2)
    b1 = g.conv(g.conv(x, "Mixed_7a_Branch_1_Conv2d_0a_1x1", same=True), "Mixed_7a_Branch_1_Conv2d_1a_3x3", stride=2)
    b2 = g.conv(g.conv(g.conv(x, "Mixed_7a_Branch_2_Conv2d_0a_1x1", same=True), "Mixed_7a_Branch_2_Conv2d_0b_3x3", same=True), "Mixed_7a_Branch_2_Conv2d_1a_3x3", stride=2)
    x = g.concat([b0, b1, b2, g.maxpool(x, "Mixed_7a_Branch_3_MaxPool_1a_3x3")], "Mixed_7a")
    for i in range(1, 7):
        p = f"Block8_{i}"
        b0 = g.conv(x, f"{p}_Branch_0_Conv2d_1x1", same=True)
        b1 = g.conv(g.conv(g.conv(x, f"{p}_Branch_1_Conv2d_0a_1x1", same=True), f"{p}_Branch_1_Conv2d_0b_1x3", same=True), f"{p}_Branch_1_Conv2d_0c_3x1", same=True)
        x = g.residual(x, [b0, b1], p, 0.2 if i < 6 else 1.0, i < 6)
    x = g.op("GlobalAveragePool", f"{prefix}/AvgPool/Mean", [x])
    x = g.op("Flatten", f"{prefix}/AvgPool/Mean_Squeeze__1234", [x], [attr_int("axis", 1)])
    w = g.const(f"{prefix}/Bottleneck/MatMul/ReadVariableOp:0", tensors["Bottleneck/kernel"])
    x = g.op("MatMul", f"{prefix}/Bottleneck/MatMul", [x, w])
    bn = "Bottleneck_BatchNorm"
    # Dense + BatchNorm on [N, D]: TensorFlow's inference form, x * (rsqrt(var + eps)) + (beta - mean * rsqrt(var + eps))
    inv = 1.0 / np.sqrt(tensors[bn + "/moving_variance"].astype(np.float64) + 1e-3)
    mul = g.op("Mul", f"{prefix}/{bn}/batchnorm/mul_1", [x, g.const(f"{prefix}/{bn}/batchnorm/mul:0", inv.astype(np.float32)[None, :])])
    off = (tensors[bn + "/beta"] - tensors[bn + "/moving_mean"] * inv).astype(np.float32)
    y = g.op("Add", f"{prefix}/{bn}/batchnorm/add_1", [mul, g.const(f"{prefix}/{bn}/batchnorm/sub:0", off[None, :])])
    g.nodes[-1] = node("Add", f"{prefix}/{bn}/batchnorm/add_1", [mul, f"{prefix}/{bn}/batchnorm/sub:0"], ["Bottleneck_BatchNorm"])
    graph = b"".join(_ld(1, n) for n in g.nodes) + _ld(2, b"tf2onnx") + b"".join(_ld(5, t) for t in g.inits) + \
        _ld(11, value_info("input_1", [1, 160, 160, 3])) + _ld(12, value_info("Bottleneck_BatchNorm", [1, D]))
    model = _vi(1, 7) + _ld(2, b"tf2onnx") + _ld(7, graph) + _ld(8, _ld(1, b"") + _vi(2, 13))
    with open(path, "wb") as f:
        f.write(model)
    return y
