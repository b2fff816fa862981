"""Execution plan for the FaceNet (Inception-ResNet-v1) conv stack on the sm_100a engine.

The reference runs this graph through onnxruntime (facenet_gpu.py:72,127); the graph itself
is the deepface/Keras Inception-ResNet-v1.  Here it is lowered to a flat
list of engine ops over NHWC fp16 activation buffers:

  * every Conv+BN(+ReLU) / biased "up" conv becomes one implicit-GEMM launch
    (M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin, K ordered tap-major/channel-minor);
  * sibling 1x1 branch heads that read the same tensor are fused horizontally into one GEMM
    (Block35: 3x(256->32) -> 256->96, Block17: 896->256, Block8: 1792->384, Mixed_7a: 896->768);
  * concat is free: producers write at a channel offset of the consumer's buffer;
  * the scaled residual `x = relu(x + s*up(cat))` (facenet_gpu.py:132-143 `scaling`) is the
    up-conv's epilogue, with s folded into its weights and bias;
  * BatchNorm (eps=1e-3, scale=False) is folded into weights/bias in fp32 before fp16 rounding.

This module is pure Python/numpy (no torch, no CUDA): it only describes work.  The engine in
csrc/facenet_engine.cu executes it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

BN_EPS = 1e-3
IN_HW = 160
IN_C_PAD = 8          # network input is NHWC fp16 with the 3 colour channels padded to 8 (one 16-byte chunk)
K_BLOCK = 64          # K elements per pipeline stage (64 fp16 = one 128-byte swizzle row)

OP_CONV, OP_MAXPOOL, OP_GAP = 1, 2, 3
F_RELU, F_RESIDUAL, F_OUT_F32 = 1, 2, 4


@dataclass
class Buf:
    H: int
    W: int
    C: int
    elt: int = 2                 # bytes per element (2 = fp16, 4 = f32)
    first: int = -1              # first op index that touches it
    last: int = -1               # last op index that touches it
    offset: int = -1             # per-image byte offset in the workspace (assigned by allocate())
    external: bool = False       # network input / outputs are bound by the caller, not the workspace
    Wp: int = 0                  # row pitch in pixels (0 = W).  The outputs of the wide stem convs are stored with pitch
                                 # W + kw - 1 so the strip kernel can use flat 128-position tiles (csrc/conv_strip.cuh)

    @property
    def bytes_per_image(self) -> int:
        return self.H * (self.Wp or self.W) * self.C * self.elt


@dataclass
class Slice:
    buf: int
    c_off: int
    c: int


@dataclass
class ConvPart:
    """One Keras conv layer contributing rows (output channels) to a (possibly fused) GEMM."""
    name: str
    cout: int
    bn: bool               # True: Conv(no bias)+BN ; False: conv with bias ("up")
    scale: float = 1.0     # residual scale folded into an "up" conv
    cin_off: int = 0       # grouped fusion: this part reads input channels [cin_off, cin_off + cin_len) of the op's
    cin_len: int = 0       # source slice (0 = all of them); its rows get zero weights for the other channels
    cin_perm: Optional[List[int]] = None   # our input channel i is Keras input channel cin_perm[i] (concat stored in another order)


@dataclass
class Op:
    kind: int
    src: Slice
    dst: Slice
    H: int = 0
    W: int = 0
    Ho: int = 0
    Wo: int = 0
    kh: int = 1
    kw: int = 1
    stride: int = 1
    pad_h: int = 0
    pad_w: int = 0
    cin: int = 0                 # logical input channels of the GEMM (padded channels included)
    cin_real: int = 0            # channels that carry weights (3 for the first conv, else == cin)
    cout: int = 0
    flags: int = 0
    res: Optional[Slice] = None
    parts: List[ConvPart] = field(default_factory=list)
    in_scale: float = 1.0        # multiplies the weights (1/255 for the first conv: pixel-scale input)
    bn_tile: int = 0             # N tile (UMMA N) the engine uses for this GEMM
    w_off: int = 0               # byte offset of the packed [cout][k_pad] fp16 weights
    b_off: int = 0               # byte offset of the fp32 bias
    label: str = ""
    s2d: bool = False            # first conv only: executed as a 2x2 stride-1 conv over the space-to-depth input (see Plan)
    pair: bool = False           # executed over PIXEL PAIRS: two horizontally adjacent pixels are one position with twice the
                                 # channels (same memory), so N doubles and the tile count halves (see Plan.pair_stem)

    @property
    def k_real(self) -> int:
        return self.kh * self.kw * self.cin

    @property
    def k_pad(self) -> int:
        return (self.k_real + K_BLOCK - 1) // K_BLOCK * K_BLOCK

    @property
    def macs_per_image(self) -> int:
        if self.kind != OP_CONV:
            return 0
        return sum(self.Ho * self.Wo * p.cout * self.kh * self.kw * (p.cin_len or self.cin_real) for p in self.parts)


def _pick_bn_tile(cout: int) -> int:
    """UMMA N for M=128 must be a multiple of 16 in [16,256]; split wide layers evenly."""
    if cout <= 256:
        return cout
    for parts in range(2, 16):
        if cout % parts == 0 and cout // parts <= 256 and (cout // parts) % 16 == 0:
            return cout // parts
    raise ValueError(f"no N tile for cout={cout}")


class Plan:
    def __init__(self, D: int, fuse_siblings: bool = True, reuse_buffers: bool = True, pitched: bool = True,
                 s2d_input: bool = True, pair_stem: bool = True):
        assert D in (128, 512)
        self.D = D
        self.pitched = pitched
        # s2d_input: the network input is handed over space-to-depth, [B, 80, 80, 16] fp16 with channel
        # (dy * 2 + dx) * 3 + c of position (Y, X) = pixel (2Y + dy, 2X + dx), channel c (12 used, 4 zero).  The
        # stride-2 3x3 first conv then IS a stride-1 2x2 conv with K = 64 over that tensor (taps (a, b) cover
        # rows 2a..2a+1, cols 2b..2b+1 of the 3x3 window; the 4th row/column gets zero weights), which the
        # strip kernel runs without im2col.  The Plan keeps describing the logical 3x3/2 conv; weights.pack()
        # writes the transformed op into the blob.
        self.s2d_input = s2d_input
        # pair_stem: Conv2d_1a and Conv2d_2a have only 32 output channels, and a tcgen05.mma costs the issuing thread the
        # same ~100 cycles at N = 32 as at N = 64.  An NHWC row of pixels with C channels IS a row of pixel pairs with 2C
        # channels, so the 'valid' kh x kw conv over pixels is a kh x 2 conv over pairs (output pair X = pixels 2X, 2X+1
        # reads input pixels 2X .. 2X+kw = input pairs X, X+1) with weights [2 Cout][kh][2][2 Cin] that are the original
        # taps shifted by the output parity (zero where the shifted tap leaves the window): a third more MMA work, a
        # third fewer MMA instructions, half as many tiles.  Row pitches must be even (80 pixels = 40 pairs).  Like s2d
        # this is a property of the packed blob: the Plan keeps describing the logical conv, weights.pack() emits the
        # pair op over alias views of the same buffers.
        self.pair_stem = pair_stem and pitched and s2d_input
        self.fuse = fuse_siblings
        self.reuse = reuse_buffers
        self.bufs: List[Buf] = []
        self.ops: List[Op] = []
        self.in_buf = self._buf(IN_HW, IN_HW, IN_C_PAD, external=True)
        self._build()
        self.allocate()

    # ---- construction helpers -------------------------------------------------------------
    def _buf(self, H, W, C, elt=2, external=False, Wp=0) -> int:
        self.bufs.append(Buf(H, W, C, elt, external=external, Wp=Wp if self.pitched else 0))
        return len(self.bufs) - 1

    def _touch(self, s: Optional[Slice], idx: int):
        if s is None:
            return
        b = self.bufs[s.buf]
        if b.first < 0:
            b.first = idx
        b.last = idx

    def _push(self, op: Op) -> Op:
        idx = len(self.ops)
        for s in (op.src, op.dst, op.res):
            self._touch(s, idx)
        self.ops.append(op)
        return op

    def whole(self, buf: int) -> Slice:
        return Slice(buf, 0, self.bufs[buf].C)

    def conv(self, parts, src: Slice, dst: Slice, kh=1, kw=1, stride=1, same=False,
             relu=True, res: Optional[Slice] = None, out_f32=False, cin_real=None,
             in_scale=1.0, label="") -> Op:
        if isinstance(parts, str):
            parts = [ConvPart(parts, dst.c, True)]
        sb, db = self.bufs[src.buf], self.bufs[dst.buf]
        H, W = sb.H, sb.W
        if same:
            assert stride == 1
            ph, pw = kh // 2, kw // 2
            Ho, Wo = H, W
        else:
            ph = pw = 0
            Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
        assert (db.H, db.W) == (Ho, Wo), (label, (db.H, db.W), (Ho, Wo))
        cout = sum(p.cout for p in parts)
        assert cout == dst.c and src.c % 8 == 0 and cout % 16 == 0
        flags = (F_RELU if relu else 0) | (F_RESIDUAL if res is not None else 0) | (F_OUT_F32 if out_f32 else 0)
        return self._push(Op(OP_CONV, src, dst, H, W, Ho, Wo, kh, kw, stride, ph, pw, src.c,
                             cin_real if cin_real is not None else src.c, cout, flags, res,
                             list(parts), in_scale, _pick_bn_tile(cout), label=label or parts[0].name))

    def maxpool(self, src: Slice, dst: Slice, label="maxpool") -> Op:
        sb, db = self.bufs[src.buf], self.bufs[dst.buf]
        Ho, Wo = (sb.H - 3) // 2 + 1, (sb.W - 3) // 2 + 1
        assert (db.H, db.W) == (Ho, Wo) and src.c == dst.c
        return self._push(Op(OP_MAXPOOL, src, dst, sb.H, sb.W, Ho, Wo, 3, 3, 2, cin=src.c, cout=src.c, label=label))

    # ---- the graph ---------------------------------------------------------
    def _build(self):
        B = self._buf
        x = self.whole(self.in_buf)
        t = B(79, 79, 32, Wp=80 if self.s2d_input else 0)
        op = self.conv("Conv2d_1a_3x3", x, self.whole(t), 3, 3, 2, cin_real=3, in_scale=1.0 / 255.0)
        op.s2d, op.pair = self.s2d_input, self.pair_stem
        x = self.whole(t)
        t = B(77, 77, 32, Wp=80 if self.pair_stem else 79)
        self.conv("Conv2d_2a_3x3", x, self.whole(t), 3, 3).pair = self.pair_stem
        x = self.whole(t)
        t = B(77, 77, 64, Wp=79); self.conv("Conv2d_2b_3x3", x, self.whole(t), 3, 3, same=True); x = self.whole(t)
        t = B(38, 38, 64); self.maxpool(x, self.whole(t), "MaxPool_3a_3x3"); x = self.whole(t)
        t = B(38, 38, 80); self.conv("Conv2d_3b_1x1", x, self.whole(t)); x = self.whole(t)
        t = B(36, 36, 192); self.conv("Conv2d_4a_3x3", x, self.whole(t), 3, 3); x = self.whole(t)
        trunk = B(17, 17, 256); self.conv("Conv2d_4b_3x3", x, self.whole(trunk), 3, 3, 2); x = self.whole(trunk)

        # same for the five Block35 blocks (csrc/block35_fused.cuh: one launch, one image per CTA pass)
        pp35 = [B(17, 17, 256), B(17, 17, 256)]
        chain35_first = len(self.ops)
        for i in range(1, 6):
            x = self._block35(x, i, pp35[(i - 1) % 2])
        for b in pp35 + [trunk]:
            self.bufs[b].first = min(self.bufs[b].first, chain35_first)
            self.bufs[b].last = len(self.ops) - 1

        # Mixed_6a: [b0 384 | b1 256 | pool 256] -> 8x8x896
        m6 = B(8, 8, 896)
        self.conv("Mixed_6a_Branch_0_Conv2d_1a_3x3", x, Slice(m6, 0, 384), 3, 3, 2)
        t1 = B(17, 17, 192); self.conv("Mixed_6a_Branch_1_Conv2d_0a_1x1", x, self.whole(t1))
        t2 = B(17, 17, 192); self.conv("Mixed_6a_Branch_1_Conv2d_0b_3x3", self.whole(t1), self.whole(t2), 3, 3, same=True)
        self.conv("Mixed_6a_Branch_1_Conv2d_1a_3x3", self.whole(t2), Slice(m6, 384, 256), 3, 3, 2)
        self.maxpool(x, Slice(m6, 640, 256), "Mixed_6a_pool")
        x = self.whole(m6)

        # The ten Block17 blocks ping-pong between two buffers that live for the whole chain, and the chain's input
        # stays intact until its last op: csrc/block17_fused.cuh runs the chain tile by tile in ONE launch (a CTA may
        # start a new tile at block 1 while others are at block 10), so no chain buffer may be recycled inside it.
        pp = [B(8, 8, 896), B(8, 8, 896)]
        chain_first = len(self.ops)
        for i in range(1, 11):
            x = self._block17(x, i, pp[(i - 1) % 2])
        for b in pp + [m6]:
            self.bufs[b].first = min(self.bufs[b].first, chain_first)
            self.bufs[b].last = len(self.ops) - 1

        # Mixed_7a: [b0 384 | b1 256 | b2 256 | pool 896] -> 3x3x1792
        m7 = B(3, 3, 1792)
        heads = ["Mixed_7a_Branch_0_Conv2d_0a_1x1", "Mixed_7a_Branch_1_Conv2d_0a_1x1", "Mixed_7a_Branch_2_Conv2d_0a_1x1"]
        if self.fuse:
            h = B(8, 8, 768)
            self.conv([ConvPart(n, 256, True) for n in heads], x, self.whole(h), label="Mixed_7a_heads")
            h0, h1, h2 = Slice(h, 0, 256), Slice(h, 256, 256), Slice(h, 512, 256)
        else:
            hs = []
            for n in heads:
                hb = B(8, 8, 256); self.conv(n, x, self.whole(hb)); hs.append(self.whole(hb))
            h0, h1, h2 = hs
        self.conv("Mixed_7a_Branch_0_Conv2d_1a_3x3", h0, Slice(m7, 0, 384), 3, 3, 2)
        self.conv("Mixed_7a_Branch_1_Conv2d_1a_3x3", h1, Slice(m7, 384, 256), 3, 3, 2)
        t = B(8, 8, 256); self.conv("Mixed_7a_Branch_2_Conv2d_0b_3x3", h2, self.whole(t), 3, 3, same=True)
        self.conv("Mixed_7a_Branch_2_Conv2d_1a_3x3", self.whole(t), Slice(m7, 640, 256), 3, 3, 2)
        self.maxpool(x, Slice(m7, 896, 896), "Mixed_7a_pool")
        x = self.whole(m7)

        for i in range(1, 6):
            x = self._block8(x, i, 0.2, True)
        x = self._block8(x, 6, 1.0, False)

        # tail: GAP -> Dense(no bias)+BN as a 1x1 "conv" over a 1x1 image, fp32 out
        g = B(1, 1, 1792)
        self._push(Op(OP_GAP, x, self.whole(g), 3, 3, 1, 1, cin=1792, cout=1792, label="AvgPool"))
        self.out_buf = B(1, 1, self.D, elt=4, external=True)
        self.conv([ConvPart("Bottleneck", self.D, True)], self.whole(g), self.whole(self.out_buf),
                  relu=False, out_f32=True, label="Bottleneck")

    def _block35(self, x: Slice, i: int, y: int) -> Slice:
        p = f"Block35_{i}"
        B = self._buf
        if self.fuse:
            # X = [b1_mid 32 | b2_mid 32 | b0 32 | b2_out 32 | b1_out 32 | b2_t 32]
            #   heads          -> X[0:96]                     (one 256->96 GEMM)
            #   3x3 (grouped)  :  X[0:64]  -> X[128:192]      (Branch_1 0b and Branch_2 0b in ONE launch: block-diagonal weights)
            #   Branch_2 0c    :  X[160:192] -> X[96:128]
            #   up reads the concat X[64:160] = [b0 | b2_out | b1_out] (Keras order is b0, b1, b2: channel permutation folded
            #   into the up conv's weights)
            X = B(17, 17, 192)
            self.conv([ConvPart(f"{p}_Branch_1_Conv2d_0a_1x1", 32, True),
                       ConvPart(f"{p}_Branch_2_Conv2d_0a_1x1", 32, True),
                       ConvPart(f"{p}_Branch_0_Conv2d_1x1", 32, True)], x, Slice(X, 0, 96), label=f"{p}_heads")
            self.conv([ConvPart(f"{p}_Branch_1_Conv2d_0b_3x3", 32, True, cin_off=0, cin_len=32),
                       ConvPart(f"{p}_Branch_2_Conv2d_0b_3x3", 32, True, cin_off=32, cin_len=32)],
                      Slice(X, 0, 64), Slice(X, 128, 64), 3, 3, same=True, label=f"{p}_Branch_12_Conv2d_0b_3x3")
            self.conv(f"{p}_Branch_2_Conv2d_0c_3x3", Slice(X, 160, 32), Slice(X, 96, 32), 3, 3, same=True)
            perm = list(range(0, 32)) + list(range(64, 96)) + list(range(32, 64))
            self.conv([ConvPart(f"{p}_Conv2d_1x1", 256, False, 0.17, cin_perm=perm)], Slice(X, 64, 96), self.whole(y), relu=True, res=x,
                      label=f"{p}_up")
            return self.whole(y)
        X = B(17, 17, 96)
        self.conv(f"{p}_Branch_0_Conv2d_1x1", x, Slice(X, 0, 32))
        t1 = B(17, 17, 32); self.conv(f"{p}_Branch_1_Conv2d_0a_1x1", x, self.whole(t1)); b1m = self.whole(t1)
        t2 = B(17, 17, 32); self.conv(f"{p}_Branch_2_Conv2d_0a_1x1", x, self.whole(t2)); b2m = self.whole(t2)
        cat, b1o, b2o = self.whole(X), Slice(X, 32, 32), Slice(X, 64, 32)
        self.conv(f"{p}_Branch_1_Conv2d_0b_3x3", b1m, b1o, 3, 3, same=True)
        t = B(17, 17, 32); self.conv(f"{p}_Branch_2_Conv2d_0b_3x3", b2m, self.whole(t), 3, 3, same=True)
        self.conv(f"{p}_Branch_2_Conv2d_0c_3x3", self.whole(t), b2o, 3, 3, same=True)
        self.conv([ConvPart(f"{p}_Conv2d_1x1", 256, False, 0.17)], cat, self.whole(y), relu=True, res=x, label=f"{p}_up")
        return self.whole(y)

    def _block17(self, x: Slice, i: int, y: int) -> Slice:
        p = f"Block17_{i}"
        B = self._buf
        if self.fuse:
            # X = [b1_mid 128 | b0 128 | b1_out 128]; up reads X[128:384]
            X = B(8, 8, 384)
            self.conv([ConvPart(f"{p}_Branch_1_Conv2d_0a_1x1", 128, True),
                       ConvPart(f"{p}_Branch_0_Conv2d_1x1", 128, True)], x, Slice(X, 0, 256), label=f"{p}_heads")
            b1m, cat, b1o = Slice(X, 0, 128), Slice(X, 128, 256), Slice(X, 256, 128)
        else:
            X = B(8, 8, 256)
            self.conv(f"{p}_Branch_0_Conv2d_1x1", x, Slice(X, 0, 128))
            t1 = B(8, 8, 128); self.conv(f"{p}_Branch_1_Conv2d_0a_1x1", x, self.whole(t1)); b1m = self.whole(t1)
            cat, b1o = self.whole(X), Slice(X, 128, 128)
        t = B(8, 8, 128); self.conv(f"{p}_Branch_1_Conv2d_0b_1x7", b1m, self.whole(t), 1, 7, same=True)
        self.conv(f"{p}_Branch_1_Conv2d_0c_7x1", self.whole(t), b1o, 7, 1, same=True)
        self.conv([ConvPart(f"{p}_Conv2d_1x1", 896, False, 0.1)], cat, self.whole(y), relu=True, res=x, label=f"{p}_up")
        return self.whole(y)

    def _block8(self, x: Slice, i: int, scale: float, relu: bool) -> Slice:
        p = f"Block8_{i}"
        B = self._buf
        if self.fuse:
            X = B(3, 3, 576)     # [b1_mid 192 | b0 192 | b1_out 192]; up reads X[192:576]
            self.conv([ConvPart(f"{p}_Branch_1_Conv2d_0a_1x1", 192, True),
                       ConvPart(f"{p}_Branch_0_Conv2d_1x1", 192, True)], x, Slice(X, 0, 384), label=f"{p}_heads")
            b1m, cat, b1o = Slice(X, 0, 192), Slice(X, 192, 384), Slice(X, 384, 192)
        else:
            X = B(3, 3, 384)
            self.conv(f"{p}_Branch_0_Conv2d_1x1", x, Slice(X, 0, 192))
            t1 = B(3, 3, 192); self.conv(f"{p}_Branch_1_Conv2d_0a_1x1", x, self.whole(t1)); b1m = self.whole(t1)
            cat, b1o = self.whole(X), Slice(X, 192, 192)
        t = B(3, 3, 192); self.conv(f"{p}_Branch_1_Conv2d_0b_1x3", b1m, self.whole(t), 1, 3, same=True)
        self.conv(f"{p}_Branch_1_Conv2d_0c_3x1", self.whole(t), b1o, 3, 1, same=True)
        y = B(3, 3, 1792)
        self.conv([ConvPart(f"{p}_Conv2d_1x1", 1792, False, scale)], cat, self.whole(y), relu=relu, res=x, label=f"{p}_up")
        return self.whole(y)

    # ---- workspace layout ------------------------------------------------------------------
    def allocate(self, align: int = 256):
        """Greedy first-fit of per-image byte offsets using buffer live ranges, so dead
        activations are overwritten while still in L2 instead of being written back to HBM."""
        order = sorted((i for i, b in enumerate(self.bufs) if not b.external), key=lambda i: self.bufs[i].first)
        placed: List[int] = []
        top = 0
        for i in order:
            b = self.bufs[i]
            size = (b.bytes_per_image + align - 1) // align * align
            if not self.reuse:          # debug layout: every buffer keeps its own memory
                b.offset = top
                top += size
                continue
            busy = sorted((self.bufs[j].offset, self.bufs[j].offset + (self.bufs[j].bytes_per_image + align - 1) // align * align)
                          for j in placed if not (self.bufs[j].last < b.first or self.bufs[j].first > b.last))
            off = 0
            for lo, hi in busy:
                if off + size <= lo:
                    break
                off = max(off, hi)
            b.offset = off
            placed.append(i)
            top = max(top, off + size)
        self.workspace_bytes_per_image = top

    # ---- queries ---------------------------------------------------------------------------
    def conv_ops(self) -> List[Op]:
        return [o for o in self.ops if o.kind == OP_CONV]

    def macs_per_image(self) -> int:
        return sum(o.macs_per_image for o in self.ops)

    def keras_tensor_shapes(self) -> dict:
        """name -> shape of every weight tensor the plan consumes (Keras naming, HWIO kernels)."""
        s = {}
        for o in self.conv_ops():
            for p in o.parts:
                if p.name == "Bottleneck":
                    s["Bottleneck/kernel"] = (o.cin, p.cout)
                else:
                    s[p.name + "/kernel"] = (o.kh, o.kw, p.cin_len or o.cin_real, p.cout)
                if p.bn:
                    for q in ("beta", "moving_mean", "moving_variance"):
                        s[f"{p.name}_BatchNorm/{q}"] = (p.cout,)
                else:
                    s[p.name + "/bias"] = (p.cout,)
        return s
