"""FaceNet weights for the sm_100a engine: sources, BN folding, fp16 packing.

Reference: facenet_gpu.py:99-106 picks weights/facenet{128,512}.onnx; in the reference
checkout those are git-LFS pointers, so two sources exist here:

  * `load_onnx_weights(path)`  - reads a real FaceNet ONNX export (see onnx_reader.py);
  * `synthetic_weights(D, seed)` - seeded random tensors of exactly the same names/shapes
    (He-scaled so activations stay O(1) through all 132 convs); used by tests, smoke, bench.

`pack(plan, tensors)` folds BatchNorm (eps=1e-3, scale=False) and the residual scale into
weights/bias in fp32, rounds weights to fp16 (bf16 misses the 0.9999 cosine gate) in the K order the implicit-GEMM kernel gathers
(tap-major, channel-minor, K padded to 64) and emits one binary blob = header + buffer table
+ op table + weights, which `fire_facenet_create` (include/fire_b200.h) consumes.
"""
from __future__ import annotations

import numpy as np

from .netplan import BN_EPS, OP_CONV, Plan

MAGIC = b"FIREB200"
BLOB_VERSION = 5

HEADER_DT = np.dtype([("magic", "S8"), ("version", "<i4"), ("D", "<i4"), ("n_ops", "<i4"), ("n_bufs", "<i4"),
                      ("ws_bytes_per_image", "<i8"), ("weights_off", "<i8"), ("weights_bytes", "<i8"),
                      ("in_buf", "<i4"), ("out_buf", "<i4")])
BUF_DT = np.dtype([("H", "<i4"), ("W", "<i4"), ("C", "<i4"), ("elt", "<i4"), ("offset", "<i8"),
                   ("external", "<i4"), ("Wp", "<i4")])
OP_DT = np.dtype([("kind", "<i4"), ("src_buf", "<i4"), ("src_coff", "<i4"), ("dst_buf", "<i4"), ("dst_coff", "<i4"),
                  ("res_buf", "<i4"), ("res_coff", "<i4"), ("H", "<i4"), ("W", "<i4"), ("Ho", "<i4"), ("Wo", "<i4"),
                  ("kh", "<i4"), ("kw", "<i4"), ("stride", "<i4"), ("pad_h", "<i4"), ("pad_w", "<i4"),
                  ("cin", "<i4"), ("cout", "<i4"), ("k_pad", "<i4"), ("flags", "<i4"), ("bn_tile", "<i4"),
                  ("flop_k", "<i4"), ("w_off", "<i8"), ("b_off", "<i8")])


F16_MAX = 65504.0


def f32_to_f16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> IEEE half bit pattern (uint16), saturating at +-65504."""
    x = np.clip(np.ascontiguousarray(x, dtype=np.float32), -F16_MAX, F16_MAX)
    return x.astype(np.float16).view(np.uint16)


def f16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(b, dtype=np.uint16).view(np.float16).astype(np.float32)


def calibration_images(n: int = 16, seed: int = 77) -> np.ndarray:
    """Seeded uint8 [n,160,160,3] images with both smooth structure and pixel noise (so that
    different images produce clearly different embeddings)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 160, 160, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:160, 0:160].astype(np.float32) / 160.0
    for i in range(n):
        img = np.zeros((160, 160, 3), dtype=np.float32)
        for _ in range(6):                                   # a few random low-frequency waves per channel
            fx, fy, ph = rng.uniform(-4, 4), rng.uniform(-4, 4), rng.uniform(0, 6.28)
            amp = rng.uniform(10, 60, size=3).astype(np.float32)
            img += amp[None, None, :] * np.sin(6.28318 * (fx * xx + fy * yy) + ph)[..., None]
        img += rng.uniform(60, 190, size=3).astype(np.float32)[None, None, :]
        img += rng.standard_normal((160, 160, 3)).astype(np.float32) * rng.uniform(2, 25)
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def synthetic_weights(D: int = 512, seed: int = 1234, calibrate: bool = True) -> dict:
    """Seeded stand-in for weights/facenet{128,512}.onnx (same tensor names and shapes).

    Kernels are He-scaled Gaussians.  With calibrate=True (default) every BatchNorm's
    moving_mean / moving_variance are then set to the statistics that layer actually sees on
    `calibration_images()` (a data-dependent init, computed once on the CPU with torch while
    *generating* the file-equivalent tensors; it is not part of inference).  That keeps all 132
    layers centred and O(1) like a trained network, so that embeddings of different images
    differ and a parity cosine is a meaningful number.  calibrate=False gives purely analytic
    statistics (fast; used where only shapes matter).
    """
    plan = Plan(D, fuse_siblings=False)
    shapes = plan.keras_tensor_shapes()
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in shapes.items():
        if name.endswith("/kernel"):
            fan_in = int(np.prod(shape[:-1]))
            base = name[:-len("/kernel")]
            relu_after = (base + "_BatchNorm/beta") in shapes and base != "Bottleneck"
            gain = 2.0 if relu_after else 1.0
            out[name] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(gain / fan_in)))
        elif name.endswith("/moving_variance"):
            out[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        else:   # beta, moving_mean, bias
            out[name] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1))
    if calibrate:
        _calibrate_bn(plan, out)
    return out


def _calibrate_bn(plan: Plan, t: dict) -> None:
    """Walk the (unfused) plan on the CPU in fp32 and overwrite each BN's moving statistics with
    the batch statistics of its own input on the calibration images."""
    import torch
    import torch.nn.functional as F
    from .netplan import F_RELU, F_RESIDUAL, OP_GAP, OP_MAXPOOL

    x0 = torch.from_numpy(calibration_images().astype(np.float32) / np.float32(255.0))
    acts = {plan.in_buf: torch.cat([x0, torch.zeros(*x0.shape[:3], 5)], dim=3)}

    def get(s):
        return acts[s.buf][..., s.c_off:s.c_off + s.c]

    def put(s, H, W, y):
        if s.buf not in acts:
            acts[s.buf] = torch.zeros(y.shape[0], H, W, plan.bufs[s.buf].C)
        acts[s.buf][..., s.c_off:s.c_off + s.c] = y

    with torch.no_grad():
        for op in plan.ops:
            src = get(op.src)
            if op.kind == OP_MAXPOOL:
                y = F.max_pool2d(src.permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
            elif op.kind == OP_GAP:
                y = src.mean(dim=(1, 2), keepdim=True)
            else:
                (p,) = op.parts
                k = np.asarray(t[p.name + "/kernel"], dtype=np.float32)
                if k.ndim == 2:
                    k = k.reshape(1, 1, *k.shape)
                kt = torch.from_numpy(k).permute(3, 2, 0, 1).contiguous()
                y = F.conv2d(src[..., :k.shape[2]].permute(0, 3, 1, 2), kt, None, stride=op.stride,
                             padding=(op.pad_h, op.pad_w)).permute(0, 2, 3, 1)
                if p.bn:
                    mean = y.mean(dim=(0, 1, 2))
                    var = y.var(dim=(0, 1, 2), unbiased=False)
                    t[p.name + "_BatchNorm/moving_mean"] = mean.numpy().astype(np.float32)
                    t[p.name + "_BatchNorm/moving_variance"] = np.maximum(var.numpy(), 1e-4).astype(np.float32)
                    beta = torch.from_numpy(t[p.name + "_BatchNorm/beta"])
                    y = (y - mean) / torch.sqrt(torch.from_numpy(t[p.name + "_BatchNorm/moving_variance"]) + BN_EPS) + beta
                else:
                    y = (y + torch.from_numpy(t[p.name + "/bias"])) * p.scale
                if op.flags & F_RESIDUAL:
                    y = y + get(op.res)
                if op.flags & F_RELU:
                    y = F.relu(y)
            put(op.dst, op.Ho, op.Wo, y)


def fold_conv(op, tensors: dict):
    """-> (W [cout, kh*kw*cin] float32 in gather order, bias [cout] float32) with BN / residual
    scale / input scale folded in fp32 (SURVEY App. A folding rules)."""
    rows, biases = [], []
    for p in op.parts:
        k = np.asarray(tensors[p.name + "/kernel"], dtype=np.float32)
        if k.ndim == 2:                      # Dense [cin, cout] -> 1x1 HWIO
            k = k.reshape(1, 1, *k.shape)
        kh, kw, cin_real, cout = k.shape
        if p.cin_perm is not None:           # our concat order differs from Keras': gather the kernel's input channels
            k = k[:, :, np.asarray(p.cin_perm), :]
        if p.cin_len:                        # grouped fusion: this part only sees a sub-range of the op's input channels
            assert (kh, kw, cout) == (op.kh, op.kw, p.cout) and cin_real == p.cin_len, (p.name, k.shape)
            full = np.zeros((kh, kw, op.cin_real, cout), dtype=np.float32)
            full[:, :, p.cin_off:p.cin_off + p.cin_len, :] = k
            k, cin_real = full, op.cin_real
        assert (kh, kw, cout) == (op.kh, op.kw, p.cout) and cin_real == op.cin_real, (p.name, k.shape)
        if p.bn:
            var = np.asarray(tensors[p.name + "_BatchNorm/moving_variance"], dtype=np.float32)
            mean = np.asarray(tensors[p.name + "_BatchNorm/moving_mean"], dtype=np.float32)
            beta = np.asarray(tensors[p.name + "_BatchNorm/beta"], dtype=np.float32)
            inv = (1.0 / np.sqrt(var.astype(np.float64) + BN_EPS)).astype(np.float32)
            wscale, b = inv, beta - mean * inv
        else:
            wscale = np.full((cout,), p.scale, dtype=np.float32)
            b = np.asarray(tensors[p.name + "/bias"], dtype=np.float32) * np.float32(p.scale)
        kk = k * wscale[None, None, None, :] * np.float32(op.in_scale)
        if cin_real != op.cin:               # first conv: 3 colour channels padded to 8
            kk = np.concatenate([kk, np.zeros((kh, kw, op.cin - cin_real, cout), np.float32)], axis=2)
        rows.append(kk.transpose(3, 0, 1, 2).reshape(cout, kh * kw * op.cin))
        biases.append(b.astype(np.float32))
    return np.concatenate(rows, 0), np.concatenate(biases, 0)


S2D_C = 16          # channels of the space-to-depth network input: (dy, dx, c) -> (dy * 2 + dx) * 3 + c, 12 used
S2D_K = 64          # K of the equivalent 2x2 conv: tap (a, b) * 16 + channel


def s2d_weights(Wg: np.ndarray, cin: int) -> np.ndarray:
    """[cout, 3*3*cin] (tap-major, channel-minor, first conv) -> [cout, 64] for the 2x2 stride-1 conv over the
    space-to-depth input: W2[(a*2+b)*16 + (dy*2+dx)*3 + c] = W[(2a+dy), (2b+dx), c], zero where the 4x4 footprint
    leaves the 3x3 window."""
    cout = Wg.shape[0]
    w3 = Wg.reshape(cout, 3, 3, cin)
    out = np.zeros((cout, 2, 2, S2D_C), dtype=np.float32)
    for a in range(2):
        for b in range(2):
            for dy in range(2):
                for dx in range(2):
                    r, c_ = 2 * a + dy, 2 * b + dx
                    if r < 3 and c_ < 3:
                        out[:, a, b, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = w3[:, r, c_, :3]
    return out.reshape(cout, S2D_K)


def pair_weights(Wg: np.ndarray, kh: int, kw: int, cin: int) -> np.ndarray:
    """[cout, kh*kw*cin] (tap-major, channel-minor; stride 1, no padding) -> [2*cout, kh*kw2*2*cin] for the same conv over
    PIXEL PAIRS (netplan.Plan.pair_stem): output row d*cout + co is output pixel 2X + d, K index ((r*kw2 + pt)*2 + e)*cin + c
    is input pixel 2(X + pt) + e, i.e. original tap s = 2 pt + e - d (zero weights where s leaves [0, kw))."""
    cout = Wg.shape[0]
    w = Wg.reshape(cout, kh, kw, cin)
    kw2 = (kw + 2) // 2                       # pixels 2X .. 2X + kw  ->  pairs X .. X + kw2 - 1
    out = np.zeros((2, cout, kh, kw2, 2, cin), dtype=np.float32)
    for d in range(2):
        for pt in range(kw2):
            for e in range(2):
                s_ = 2 * pt + e - d
                if 0 <= s_ < kw:
                    out[d, :, :, pt, e, :] = w[:, :, s_, :]
    return out.reshape(2 * cout, kh * kw2 * 2 * cin)


def unpair_weights(Wpair: np.ndarray, kh: int, kw: int, cin: int) -> np.ndarray:
    """Inverse of pair_weights (test tool: the plan emulator runs the logical conv); checks that the two parity blocks
    are shifted copies of the same taps."""
    cout = Wpair.shape[0] // 2
    kw2 = (kw + 2) // 2
    v = Wpair[:, :kh * kw2 * 2 * cin].reshape(2, cout, kh, kw2, 2, cin)
    w = np.zeros((cout, kh, kw, cin), dtype=np.float32)
    for s_ in range(kw):
        w[:, :, s_, :] = v[0, :, :, s_ // 2, s_ % 2, :]
    assert np.array_equal(pair_weights(w.reshape(cout, -1), kh, kw, cin), Wpair[:, :kh * kw2 * 2 * cin].astype(np.float32))
    return w.reshape(cout, kh * kw * cin)


def pack(plan: Plan, tensors: dict) -> bytes:
    """Serialise plan + folded bf16 weights into the blob `fire_facenet_create` parses."""
    chunks, pos = [], 0

    def put(arr: np.ndarray) -> int:
        nonlocal pos
        off = pos
        raw = arr.tobytes()
        padn = (-len(raw)) % 256
        chunks.append(raw + b"\0" * padn)
        pos += len(raw) + padn
        return off

    for op in plan.ops:
        if op.kind != OP_CONV:
            continue
        W, b = fold_conv(op, tensors)
        if op.s2d:
            W = s2d_weights(W, op.cin)
        k_pad = S2D_K if op.s2d else op.k_pad
        cout = op.cout
        if op.pair:                                        # the same conv over pixel pairs: [2 cout][kh * 2 * 2 cin]
            kh, kw, cin = (2, 2, S2D_C) if op.s2d else (op.kh, op.kw, op.cin)
            assert op.stride == (2 if op.s2d else 1) and op.pad_h == 0 and op.pad_w == 0 and kw in (2, 3)
            W = pair_weights(W[:, :kh * kw * cin], kh, kw, cin)
            b = np.concatenate([b, b])
            cout, k_pad = 2 * cout, (W.shape[1] + 63) // 64 * 64
        Wp = np.zeros((cout, k_pad), dtype=np.uint16)
        Wp[:, :W.shape[1]] = f32_to_f16_bits(W)
        op.w_off = put(Wp)
        op.b_off = put(b)

    s2d = any(o.s2d for o in plan.ops)
    buf_rows = []
    for i, b in enumerate(plan.bufs):
        if s2d and i == plan.in_buf:                       # the engine sees the space-to-depth tensor
            buf_rows.append((b.H // 2, b.W // 2, S2D_C, b.elt, 0, 1, 0))
            continue
        buf_rows.append((b.H, b.W, b.C, b.elt, max(b.offset, 0), int(b.external), b.Wp))

    def pair_view(i: int, w_pairs: int) -> int:
        """Extra buffer-table entry: buffer i seen as pixel pairs (same memory; external = 2 marks a view of the network input)."""
        H, Wd, C, elt, off, ext, Wpitch = buf_rows[i]
        pitch = Wpitch or Wd
        assert pitch % 2 == 0 and w_pairs <= pitch // 2, (i, pitch, w_pairs)
        buf_rows.append((H, w_pairs, 2 * C, elt, off, 2 if ext else 0, pitch // 2))
        return len(buf_rows) - 1

    ops = np.zeros(len(plan.ops), dtype=OP_DT)
    for i, o in enumerate(plan.ops):
        res_buf, res_coff = (o.res.buf, o.res.c_off) if o.res is not None else (-1, 0)
        flop_k = o.macs_per_image // (o.Ho * o.Wo * o.cout) if o.kind == OP_CONV else 0      # real MACs per output element
        if o.pair:
            # kh x kw over [H, W, C]  ==  kh x 2 over the pair views [H, W/2, 2C]; Wo pairs = input pairs - 1, the last
            # input pair column is the flat tiles' spare column, whose first pixel is still a correct output
            H, Wd, kh, cin = (o.H // 2, o.W // 2, 2, S2D_C) if o.s2d else (o.H, plan.bufs[o.src.buf].Wp or o.W, o.kh, o.cin)
            assert o.src.c_off == 0 and o.dst.c_off == 0 and o.res is None and Wd % 2 == 0
            wp_in = Wd // 2
            src, dst = pair_view(o.src.buf, wp_in), pair_view(o.dst.buf, wp_in - 1)
            assert buf_rows[dst][6] == wp_in, "pair conv: the destination pitch must equal the input width in pairs"
            k_pad = (kh * 2 * 2 * cin + 63) // 64 * 64
            flop_k = int(round(o.macs_per_image / (o.Ho * (wp_in - 1) * 2 * o.cout)))
            ops[i] = (o.kind, src, 0, dst, 0, -1, 0, H, wp_in, o.Ho, wp_in - 1, kh, 2, 1, 0, 0, 2 * cin, 2 * o.cout, k_pad,
                      o.flags, min(2 * o.cout, 256), flop_k, o.w_off, o.b_off)
            continue
        if o.s2d:                                          # 3x3 / stride 2 over [160,160,8]  ==  2x2 / stride 1 over [80,80,16]
            ops[i] = (o.kind, o.src.buf, 0, o.dst.buf, o.dst.c_off, res_buf, res_coff, o.H // 2, o.W // 2, o.Ho, o.Wo,
                      2, 2, 1, 0, 0, S2D_C, o.cout, S2D_K, o.flags, o.bn_tile, flop_k, o.w_off, o.b_off)
            continue
        ops[i] = (o.kind, o.src.buf, o.src.c_off, o.dst.buf, o.dst.c_off, res_buf, res_coff, o.H, o.W, o.Ho, o.Wo,
                  o.kh, o.kw, o.stride, o.pad_h, o.pad_w, o.cin, o.cout, o.k_pad if o.kind == OP_CONV else 0,
                  o.flags, o.bn_tile, flop_k, o.w_off, o.b_off)
    bufs = np.zeros(len(buf_rows), dtype=BUF_DT)
    for i, row in enumerate(buf_rows):
        bufs[i] = row
    hdr = np.zeros(1, dtype=HEADER_DT)
    meta = HEADER_DT.itemsize + bufs.nbytes + ops.nbytes
    weights_off = (meta + 255) // 256 * 256
    hdr[0] = (MAGIC, BLOB_VERSION, plan.D, len(plan.ops), len(buf_rows), plan.workspace_bytes_per_image,
              weights_off, pos, plan.in_buf, plan.out_buf)
    head = hdr.tobytes() + bufs.tobytes() + ops.tobytes()
    return head + b"\0" * (weights_off - len(head)) + b"".join(chunks)
