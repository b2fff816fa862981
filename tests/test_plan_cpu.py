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


def test_pair_weights_compute_the_same_convolution():
    """netplan.Plan.pair_stem: a 'valid' kh x kw conv over [H, W, C] equals the kh x 2 conv with weights.pair_weights over the
    same memory seen as pixel pairs [H, W/2, 2C] (what Conv2d_1a / 2a run as on the GPU); unpair_weights inverts it."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    for kh, kw, cin, cout, H, Wd in ((3, 3, 8, 16, 7, 10), (2, 2, 16, 32, 6, 12)):
        x = rng.standard_normal((2, H, Wd, cin)).astype(np.float32)
        w = rng.standard_normal((cout, kh, kw, cin)).astype(np.float32)
        y = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(w).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
        wp = W.pair_weights(w.reshape(cout, -1), kh, kw, cin)
        assert np.array_equal(W.unpair_weights(wp, kh, kw, cin), w.reshape(cout, -1))
        xp = torch.from_numpy(x.reshape(2, H, Wd // 2, 2 * cin)).permute(0, 3, 1, 2)
        kp = torch.from_numpy(wp.reshape(2 * cout, kh, 2, 2 * cin)).permute(0, 3, 1, 2)
        yp = F.conv2d(xp, kp).permute(0, 2, 3, 1).numpy()              # [2, Ho, W/2 - 1, 2 cout]
        yp = yp.reshape(2, yp.shape[1], -1, cout)                       # pixels 0 .. W - 3
        n = min(y.shape[2], yp.shape[2])
        assert n >= Wd - kw - 1 and np.allclose(y[:, :, :n], yp[:, :, :n], atol=1e-4)


def test_pair_stem_blob_views():
    p = Plan(512)
    assert p.pair_stem and [o.pair for o in p.conv_ops()[:3]] == [True, True, False]
    blob = W.pack(p, W.synthetic_weights(512, 5, calibrate=False))
    hdr = np.frombuffer(blob, dtype=W.HEADER_DT, count=1)[0]
    assert hdr["n_bufs"] == len(p.bufs) + 4                          # two pair views per pair op
    bufs = np.frombuffer(blob, dtype=W.BUF_DT, count=int(hdr["n_bufs"]), offset=W.HEADER_DT.itemsize)
    ops = np.frombuffer(blob, dtype=W.OP_DT, count=len(p.ops), offset=W.HEADER_DT.itemsize + bufs.nbytes)
    for i in (0, 1):
        o, s, d = ops[i], bufs[ops[i]["src_buf"]], bufs[ops[i]["dst_buf"]]
        assert (o["kw"], o["cout"], o["W"], o["Wo"], d["Wp"], s["W"]) == (2, 64, 40, 39, 40, 40)
        assert d["offset"] == bufs[p.ops[i].dst.buf]["offset"] and s["C"] == o["cin"] and d["C"] == 64
    assert bufs[ops[0]["src_buf"]]["external"] == 2 and not Plan(512, pitched=False).pair_stem


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
