"""The engine's plan (fusion, channel slices, buffer reuse, BN folding, fp16 packing) emulated on the CPU must equal
the oracle graph; this is the no-GPU half of the FaceNet parity argument (the GPU half is tests/test_gpu_facenet.py)."""
import numpy as np
import pytest

import plan_emu
from fire_b200 import weights as W
from fire_b200.netplan import OP_CONV, Plan
from oracle.facenet_ref import facenet_forward, weight_shapes


def _cos(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


@pytest.fixture(scope="module")
def setup():
    t = W.synthetic_weights(512, 1234)
    rng = np.random.default_rng(2)
    u8 = np.concatenate([rng.integers(0, 256, (3, 160, 160, 3), dtype=np.uint8), W.calibration_images(3, seed=5)])
    x = u8.astype(np.float32) / 255.0
    return t, x, facenet_forward(t, x)


def test_macs_and_tensor_names_match_the_oracle_graph():
    for D, macs in ((128, 1_416_974_176), (512, 1_417_662_304)):      # SURVEY App. A totals
        for fuse in (True, False):
            p = Plan(D, fuse_siblings=fuse)
            assert p.macs_per_image() == macs
            assert p.keras_tensor_shapes() == weight_shapes(D)
    assert len(Plan(512, fuse_siblings=False).conv_ops()) == 133 and len(Plan(512).conv_ops()) == 100


def test_buffer_reuse_never_overlaps_live_buffers():
    p = Plan(512)
    live = [(i, b) for i, b in enumerate(p.bufs) if not b.external]
    for i, a in live:
        for j, b in live:
            if i < j and not (a.last < b.first or b.last < a.first):
                assert a.offset + a.bytes_per_image <= b.offset or b.offset + b.bytes_per_image <= a.offset, (i, j)
    assert p.workspace_bytes_per_image < 2_000_000                   # ~1.5 MB/image instead of 6.5 MB without reuse


@pytest.mark.parametrize("fuse", [True, False])
def test_plan_equals_oracle_fp32_storage(setup, fuse):
    t, x, ref = setup
    p = Plan(512, fuse_siblings=fuse)
    out = plan_emu.run_plan(p, W.pack(p, t), plan_emu.to_pixel_nhwc8(x), store_f16=False)
    assert _cos(out, ref).min() > 0.99995                             # only the fp16 weight rounding separates them


def test_plan_with_fp16_storage_meets_the_parity_bar(setup):
    t, x, ref = setup
    p = Plan(512)
    out = plan_emu.run_plan(p, W.pack(p, t), plan_emu.to_pixel_nhwc8(x), store_f16=True)
    assert _cos(out, ref).min() >= 0.9999


def test_bf16_would_miss_the_bar(setup, monkeypatch):
    """Why the engine computes in fp16 and not bf16: same plan, bf16 weights + activations."""
    import torch
    t, x, ref = setup
    monkeypatch.setattr(W, "f32_to_f16_bits", lambda a: (np.ascontiguousarray(a, np.float32).view(np.uint32) + 0x8000 >> 16).astype(np.uint16))
    monkeypatch.setattr(W, "f16_bits_to_f32", lambda b: (np.ascontiguousarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32))
    monkeypatch.setattr(plan_emu, "_f16", lambda v: v.to(torch.bfloat16).to(torch.float32))
    p = Plan(512)
    out = plan_emu.run_plan(p, W.pack(p, t), plan_emu.to_pixel_nhwc8(x), store_f16=True)
    assert _cos(out, ref).min() < 0.9999


def test_blob_layout_and_tiling_constraints():
    p = Plan(128)
    blob = W.pack(p, W.synthetic_weights(128, 5, calibrate=False))
    hdr = np.frombuffer(blob, dtype=W.HEADER_DT, count=1)[0]
    assert hdr["magic"] == W.MAGIC and hdr["n_ops"] == len(p.ops) and hdr["D"] == 128
    assert W.HEADER_DT.itemsize == 56 and W.BUF_DT.itemsize == 32 and W.OP_DT.itemsize == 104   # C structs in facenet_engine.cu
    for o in p.conv_ops():
        assert o.cin % 8 == 0 and o.cout % o.bn_tile == 0 and o.bn_tile % 16 == 0 and 16 <= o.bn_tile <= 256
        assert o.k_pad % 64 == 0 and o.w_off % 256 == 0 and o.b_off % 256 == 0
        assert o.src.c_off % 8 == 0 and o.dst.c_off % 8 == 0
