"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the arithmetic behind the reference's
``Encoder.encode -> FaceNetClient.__call__ -> onnxruntime.InferenceSession.run``
(reference modules/encoder.py:16-17, facenet_gpu.py:116-129).  The model file
(weights/facenet{128,512}.onnx, a git-LFS pointer in the reference checkout) is the
deepface/Keras Inception-ResNet-v1 that facenet_gpu.py:132-143 (`scaling`) and
lisences/NOTICE.md:13-14 point to; this file restates that published graph
 in plain torch fp32/fp64 on the CPU.

PARITY UNPINNED: the reference ships no tests, no golden vectors and no usable
weights, and onnxruntime is not installable here, so this oracle cannot be pinned
against reference outputs.  It is pinned only structurally (parameter count equals
the byte size recorded in the LFS pointers to 0.05 %, tests/test_oracle_facenet.py).

Input  : NHWC float32 [B,160,160,3] exactly as modules/encoder.py:19-27 produces it.
Output : [B,D] float32, NOT L2-normalised (the caller normalises,
         modules/face_recognition.py:225-229).

Weights come in as a dict of numpy arrays with Keras names:
  "<conv>/kernel"  [kh,kw,Cin,Cout]      (HWIO)
  "<conv>/bias"    [Cout]                 (only the residual "up" 1x1 convs)
  "<conv>_BatchNorm/{beta,moving_mean,moving_variance}" [Cout]   (scale=False, eps=1e-3)
  "Bottleneck/kernel" [1792,D], "Bottleneck_BatchNorm/{beta,moving_mean,moving_variance}" [D]
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


class _Net:
    def __init__(self, weights: dict, dtype=torch.float32, act_round=None):
        self.w = weights
        self.dtype = dtype
        # act_round: optional callable applied to every stored activation (used by
        # tests to model reduced-precision storage; None = pure fp32/fp64 oracle).
        self.act_round = act_round
        self._cache = {}

    def _t(self, name):
        """Weight tensor, converted once per net (an onnxruntime session also holds its weights resident)."""
        t = self._cache.get(name)
        if t is None:
            t = torch.from_numpy(np.ascontiguousarray(self.w[name])).to(self.dtype)
            if name.endswith("/kernel") and t.dim() == 4:
                t = t.permute(3, 2, 0, 1).contiguous()            # HWIO -> OIHW
            self._cache[name] = t
        return t

    def _r(self, x):
        return self.act_round(x) if self.act_round is not None else x

    def conv_bn_relu(self, x, name, stride=1, padding="valid"):
        """Conv2D(no bias) -> BatchNorm(eps=1e-3, scale=False) -> ReLU (App. A conventions)."""
        k = self._t(name + "/kernel")
        kh, kw = k.shape[2], k.shape[3]
        if padding == "same":
            assert stride == 1
            pad = (kh // 2, kw // 2)
        else:
            pad = (0, 0)
        y = F.conv2d(x, k, None, stride=stride, padding=pad)
        beta = self._t(name + "_BatchNorm/beta").view(1, -1, 1, 1)
        mean = self._t(name + "_BatchNorm/moving_mean").view(1, -1, 1, 1)
        var = self._t(name + "_BatchNorm/moving_variance").view(1, -1, 1, 1)
        y = (y - mean) / torch.sqrt(var + BN_EPS) + beta
        return self._r(F.relu(y))

    def up(self, x, name):
        """1x1 Conv2D with bias, no BN, no activation."""
        k = self._t(name + "/kernel")
        b = self._t(name + "/bias")
        return F.conv2d(x, k, b)

    def maxpool(self, x):
        return F.max_pool2d(x, 3, 2)

    def block35(self, x, i):
        p = f"Block35_{i}"
        b0 = self.conv_bn_relu(x, f"{p}_Branch_0_Conv2d_1x1", padding="same")
        b1 = self.conv_bn_relu(x, f"{p}_Branch_1_Conv2d_0a_1x1", padding="same")
        b1 = self.conv_bn_relu(b1, f"{p}_Branch_1_Conv2d_0b_3x3", padding="same")
        b2 = self.conv_bn_relu(x, f"{p}_Branch_2_Conv2d_0a_1x1", padding="same")
        b2 = self.conv_bn_relu(b2, f"{p}_Branch_2_Conv2d_0b_3x3", padding="same")
        b2 = self.conv_bn_relu(b2, f"{p}_Branch_2_Conv2d_0c_3x3", padding="same")
        u = self.up(torch.cat([b0, b1, b2], 1), f"{p}_Conv2d_1x1")
        return self._r(F.relu(x + 0.17 * u))          # facenet_gpu.py:132-143 `scaling`

    def block17(self, x, i):
        p = f"Block17_{i}"
        b0 = self.conv_bn_relu(x, f"{p}_Branch_0_Conv2d_1x1", padding="same")
        b1 = self.conv_bn_relu(x, f"{p}_Branch_1_Conv2d_0a_1x1", padding="same")
        b1 = self.conv_bn_relu(b1, f"{p}_Branch_1_Conv2d_0b_1x7", padding="same")
        b1 = self.conv_bn_relu(b1, f"{p}_Branch_1_Conv2d_0c_7x1", padding="same")
        u = self.up(torch.cat([b0, b1], 1), f"{p}_Conv2d_1x1")
        return self._r(F.relu(x + 0.1 * u))

    def block8(self, x, i, scale, relu):
        p = f"Block8_{i}"
        b0 = self.conv_bn_relu(x, f"{p}_Branch_0_Conv2d_1x1", padding="same")
        b1 = self.conv_bn_relu(x, f"{p}_Branch_1_Conv2d_0a_1x1", padding="same")
        b1 = self.conv_bn_relu(b1, f"{p}_Branch_1_Conv2d_0b_1x3", padding="same")
        b1 = self.conv_bn_relu(b1, f"{p}_Branch_1_Conv2d_0c_3x1", padding="same")
        u = self.up(torch.cat([b0, b1], 1), f"{p}_Conv2d_1x1")
        y = x + scale * u
        return self._r(F.relu(y) if relu else y)

    def forward(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        x = self._r(x_nhwc.to(self.dtype).permute(0, 3, 1, 2).contiguous())
        x = self.conv_bn_relu(x, "Conv2d_1a_3x3", stride=2)
        x = self.conv_bn_relu(x, "Conv2d_2a_3x3")
        x = self.conv_bn_relu(x, "Conv2d_2b_3x3", padding="same")
        x = self.maxpool(x)
        x = self.conv_bn_relu(x, "Conv2d_3b_1x1")
        x = self.conv_bn_relu(x, "Conv2d_4a_3x3")
        x = self.conv_bn_relu(x, "Conv2d_4b_3x3", stride=2)
        for i in range(1, 6):
            x = self.block35(x, i)
        # Mixed_6a
        b0 = self.conv_bn_relu(x, "Mixed_6a_Branch_0_Conv2d_1a_3x3", stride=2)
        b1 = self.conv_bn_relu(x, "Mixed_6a_Branch_1_Conv2d_0a_1x1", padding="same")
        b1 = self.conv_bn_relu(b1, "Mixed_6a_Branch_1_Conv2d_0b_3x3", padding="same")
        b1 = self.conv_bn_relu(b1, "Mixed_6a_Branch_1_Conv2d_1a_3x3", stride=2)
        x = torch.cat([b0, b1, self.maxpool(x)], 1)
        for i in range(1, 11):
            x = self.block17(x, i)
        # Mixed_7a
        b0 = self.conv_bn_relu(x, "Mixed_7a_Branch_0_Conv2d_0a_1x1", padding="same")
        b0 = self.conv_bn_relu(b0, "Mixed_7a_Branch_0_Conv2d_1a_3x3", stride=2)
        b1 = self.conv_bn_relu(x, "Mixed_7a_Branch_1_Conv2d_0a_1x1", padding="same")
        b1 = self.conv_bn_relu(b1, "Mixed_7a_Branch_1_Conv2d_1a_3x3", stride=2)
        b2 = self.conv_bn_relu(x, "Mixed_7a_Branch_2_Conv2d_0a_1x1", padding="same")
        b2 = self.conv_bn_relu(b2, "Mixed_7a_Branch_2_Conv2d_0b_3x3", padding="same")
        b2 = self.conv_bn_relu(b2, "Mixed_7a_Branch_2_Conv2d_1a_3x3", stride=2)
        x = torch.cat([b0, b1, b2, self.maxpool(x)], 1)
        for i in range(1, 6):
            x = self.block8(x, i, 0.2, True)
        x = self.block8(x, 6, 1.0, False)
        x = self._r(x.mean(dim=(2, 3)))                       # GlobalAveragePooling2D; Dropout = identity
        y = x @ self._t("Bottleneck/kernel")                  # Dense, no bias
        beta = self._t("Bottleneck_BatchNorm/beta")
        mean = self._t("Bottleneck_BatchNorm/moving_mean")
        var = self._t("Bottleneck_BatchNorm/moving_variance")
        return (y - mean) / torch.sqrt(var + BN_EPS) + beta


class FaceNetRef:
    """Session-like wrapper: weights converted once, then `__call__(x_nhwc)` like FaceNetClient.__call__."""

    def __init__(self, weights: dict, dtype=torch.float32):
        self.net = _Net(weights, dtype=dtype)

    def __call__(self, x_nhwc: np.ndarray) -> np.ndarray:
        with torch.no_grad():
            return self.net.forward(torch.from_numpy(np.ascontiguousarray(x_nhwc))).numpy()


def facenet_forward(weights: dict, x_nhwc: np.ndarray, dtype=torch.float32, act_round=None,
                    chunk: int = 16) -> np.ndarray:
    """Run the oracle on an NHWC float batch; returns [B,D] float32 (or float64)."""
    net = _Net(weights, dtype=dtype, act_round=act_round)
    outs = []
    with torch.no_grad():
        for i in range(0, x_nhwc.shape[0], chunk):
            xb = torch.from_numpy(np.ascontiguousarray(x_nhwc[i:i + chunk]))
            outs.append(net.forward(xb))
    return torch.cat(outs, 0).numpy()


def l2_normalize_rows(e: np.ndarray) -> np.ndarray:
    """modules/face_recognition.py:225-229: e / ||e||_2 (rows with zero norm are left as-is;
    the reference skips such faces)."""
    n = np.linalg.norm(e, axis=-1, keepdims=True)
    return np.where(n > 0, e / np.where(n > 0, n, 1), e)


def expected_param_count(D: int) -> int:
    """Parameter count the LFS pointer sizes pin: 22 808 144 / 23 497 424."""
    from collections import OrderedDict  # noqa: F401
    total = 0
    for name, shape in weight_shapes(D).items():
        total += int(np.prod(shape))
    return total


def weight_shapes(D: int) -> dict:
    """Names and shapes of every tensor the graph uses, derived independently of the product's plan."""
    s = {}

    def conv(name, kh, kw, cin, cout):
        s[name + "/kernel"] = (kh, kw, cin, cout)
        for p in ("beta", "moving_mean", "moving_variance"):
            s[f"{name}_BatchNorm/{p}"] = (cout,)

    def up(name, cin, cout):
        s[name + "/kernel"] = (1, 1, cin, cout)
        s[name + "/bias"] = (cout,)

    conv("Conv2d_1a_3x3", 3, 3, 3, 32)
    conv("Conv2d_2a_3x3", 3, 3, 32, 32)
    conv("Conv2d_2b_3x3", 3, 3, 32, 64)
    conv("Conv2d_3b_1x1", 1, 1, 64, 80)
    conv("Conv2d_4a_3x3", 3, 3, 80, 192)
    conv("Conv2d_4b_3x3", 3, 3, 192, 256)
    for i in range(1, 6):
        p = f"Block35_{i}"
        conv(f"{p}_Branch_0_Conv2d_1x1", 1, 1, 256, 32)
        conv(f"{p}_Branch_1_Conv2d_0a_1x1", 1, 1, 256, 32)
        conv(f"{p}_Branch_1_Conv2d_0b_3x3", 3, 3, 32, 32)
        conv(f"{p}_Branch_2_Conv2d_0a_1x1", 1, 1, 256, 32)
        conv(f"{p}_Branch_2_Conv2d_0b_3x3", 3, 3, 32, 32)
        conv(f"{p}_Branch_2_Conv2d_0c_3x3", 3, 3, 32, 32)
        up(f"{p}_Conv2d_1x1", 96, 256)
    conv("Mixed_6a_Branch_0_Conv2d_1a_3x3", 3, 3, 256, 384)
    conv("Mixed_6a_Branch_1_Conv2d_0a_1x1", 1, 1, 256, 192)
    conv("Mixed_6a_Branch_1_Conv2d_0b_3x3", 3, 3, 192, 192)
    conv("Mixed_6a_Branch_1_Conv2d_1a_3x3", 3, 3, 192, 256)
    for i in range(1, 11):
        p = f"Block17_{i}"
        conv(f"{p}_Branch_0_Conv2d_1x1", 1, 1, 896, 128)
        conv(f"{p}_Branch_1_Conv2d_0a_1x1", 1, 1, 896, 128)
        conv(f"{p}_Branch_1_Conv2d_0b_1x7", 1, 7, 128, 128)
        conv(f"{p}_Branch_1_Conv2d_0c_7x1", 7, 1, 128, 128)
        up(f"{p}_Conv2d_1x1", 256, 896)
    conv("Mixed_7a_Branch_0_Conv2d_0a_1x1", 1, 1, 896, 256)
    conv("Mixed_7a_Branch_0_Conv2d_1a_3x3", 3, 3, 256, 384)
    conv("Mixed_7a_Branch_1_Conv2d_0a_1x1", 1, 1, 896, 256)
    conv("Mixed_7a_Branch_1_Conv2d_1a_3x3", 3, 3, 256, 256)
    conv("Mixed_7a_Branch_2_Conv2d_0a_1x1", 1, 1, 896, 256)
    conv("Mixed_7a_Branch_2_Conv2d_0b_3x3", 3, 3, 256, 256)
    conv("Mixed_7a_Branch_2_Conv2d_1a_3x3", 3, 3, 256, 256)
    for i in range(1, 7):
        p = f"Block8_{i}"
        conv(f"{p}_Branch_0_Conv2d_1x1", 1, 1, 1792, 192)
        conv(f"{p}_Branch_1_Conv2d_0a_1x1", 1, 1, 1792, 192)
        conv(f"{p}_Branch_1_Conv2d_0b_1x3", 1, 3, 192, 192)
        conv(f"{p}_Branch_1_Conv2d_0c_3x1", 3, 1, 192, 192)
        up(f"{p}_Conv2d_1x1", 384, 1792)
    s["Bottleneck/kernel"] = (1792, D)
    for p in ("beta", "moving_mean", "moving_variance"):
        s[f"Bottleneck_BatchNorm/{p}"] = (D,)
    return s
