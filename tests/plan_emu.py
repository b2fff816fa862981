"""CPU emulation of the engine's *plan* (test tool, not product code, not an oracle).

Executes fire_b200.netplan.Plan op by op with torch on the CPU, using the *packed* blob
weights (BN folded, fp16-rounded, gather-ordered) and rounding every stored activation the way
the sm_100a engine does (fp16 storage, fp32 accumulate).  It answers, without a GPU:
  * is the plan (fusion, channel slices, buffer reuse, folding, packing) equivalent to the
    oracle graph?   (tests/test_plan_cpu.py compares it with oracle/facenet_ref.py)
  * what embedding error does 16-bit storage cost?  (DESIGN.md precision table)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from fire_b200 import weights as W
from fire_b200.netplan import F_OUT_F32, F_RELU, F_RESIDUAL, OP_CONV, OP_GAP, OP_MAXPOOL, Plan


def _f16(x: torch.Tensor) -> torch.Tensor:
    return x.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)


def run_plan(plan: Plan, blob: bytes, x_pix_nhwc8: np.ndarray, store_f16: bool = True,
             honor_offsets: bool = True, return_buffers: bool = False):
    """x_pix_nhwc8: [B,160,160,8] float32 pixel-scale input (channels 3..7 zero).
    honor_offsets=True stores activations in one flat per-batch arena at the plan's offsets, so
    an allocator bug (overlapping live buffers) corrupts results here exactly as it would on GPU."""
    hdr = np.frombuffer(blob, dtype=W.HEADER_DT, count=1)[0]
    wbase = int(hdr["weights_off"])
    B = x_pix_nhwc8.shape[0]
    rnd = _f16 if store_f16 else (lambda t: t)

    arena = torch.zeros(plan.workspace_bytes_per_image * B // 2 + 16, dtype=torch.float32)  # one float per fp16 slot
    ext = {}

    def view(bi: int) -> torch.Tensor:
        b = plan.bufs[bi]
        if b.external or not honor_offsets:
            if bi not in ext:
                ext[bi] = torch.zeros(B, b.H, b.W, b.C)
            return ext[bi]
        assert b.elt == 2
        start = b.offset * B // 2
        n = B * b.H * b.W * b.C
        return arena[start:start + n].view(B, b.H, b.W, b.C)

    ext[plan.in_buf] = rnd(torch.from_numpy(np.ascontiguousarray(x_pix_nhwc8)).float())
    with torch.no_grad():
        for op in plan.ops:
            src = view(op.src.buf)[..., op.src.c_off:op.src.c_off + op.src.c]
            dstv = view(op.dst.buf)
            if op.kind == OP_CONV:
                def packed(k_real, k_pad):
                    """this op's [cout][k_real] weights out of the blob; a pair op (netplan.Plan.pair_stem) is stored as
                    the pixel-pair conv [2 cout][kh * 2 * 2 cin]: undo, the emulator runs the logical conv"""
                    if not op.pair:
                        wb = np.frombuffer(blob, dtype=np.uint16, count=op.cout * k_pad, offset=wbase + op.w_off)
                        return W.f16_bits_to_f32(wb.reshape(op.cout, k_pad)[:, :k_real])
                    kh, kw, cin = (2, 2, W.S2D_C) if op.s2d else (op.kh, op.kw, op.cin)
                    kp = (kh * 2 * 2 * cin + 63) // 64 * 64
                    wb = np.frombuffer(blob, dtype=np.uint16, count=2 * op.cout * kp, offset=wbase + op.w_off)
                    return W.unpair_weights(W.f16_bits_to_f32(wb.reshape(2 * op.cout, kp)), kh, kw, cin)
                if op.s2d:      # packed as the 2x2 conv over the space-to-depth input: undo, the emulator feeds NHWC8
                    w2 = packed(W.S2D_K, W.S2D_K).reshape(op.cout, 2, 2, W.S2D_C)
                    w3 = np.zeros((op.cout, 3, 3, op.cin), dtype=np.float32)
                    for a in range(2):
                        for b_ in range(2):
                            for dy in range(2):
                                for dx in range(2):
                                    if 2 * a + dy < 3 and 2 * b_ + dx < 3:
                                        w3[:, 2 * a + dy, 2 * b_ + dx, :3] = w2[:, a, b_, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3]
                    wf = w3.reshape(op.cout, -1)
                else:
                    wf = packed(op.k_real, op.k_pad)
                bias = np.frombuffer(blob, dtype=np.float32, count=op.cout, offset=wbase + op.b_off)
                k = torch.from_numpy(wf.reshape(op.cout, op.kh, op.kw, op.cin).transpose(0, 3, 1, 2).copy())
                y = F.conv2d(src.permute(0, 3, 1, 2), k, torch.from_numpy(bias.copy()), stride=op.stride,
                             padding=(op.pad_h, op.pad_w)).permute(0, 2, 3, 1)
                if op.flags & F_RESIDUAL:
                    y = y + view(op.res.buf)[..., op.res.c_off:op.res.c_off + op.res.c]
                if op.flags & F_RELU:
                    y = F.relu(y)
                if not (op.flags & F_OUT_F32):
                    y = rnd(y)
            elif op.kind == OP_MAXPOOL:
                y = F.max_pool2d(src.permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
            elif op.kind == OP_GAP:
                y = rnd(src.mean(dim=(1, 2), keepdim=True))
            else:
                raise ValueError(op.kind)
            dstv[..., op.dst.c_off:op.dst.c_off + op.dst.c] = y
    out = view(plan.out_buf).reshape(B, plan.D).numpy().copy()
    if return_buffers:
        return out, {i: view(i).numpy().copy() for i in range(len(plan.bufs)) if plan.bufs[i].first >= 0 or i == plan.in_buf}
    return out


def to_pixel_nhwc8(x_unit_nhwc3: np.ndarray) -> np.ndarray:
    """[B,160,160,3] float in [0,1] (modules/encoder.py:21) -> [B,160,160,8] pixel-scale float."""
    B = x_unit_nhwc3.shape[0]
    out = np.zeros((B, 160, 160, 8), dtype=np.float32)
    out[..., :3] = x_unit_nhwc3 * np.float32(255.0)
    return out
