"""GPU parity: FaceNet128/512 on the tcgen05 engine (through the C ABI) vs the fp32 CPU oracle.
Tolerance (north star): cosine >= 0.9999 per image against the fp32 reference output."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cos(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


def _images(n, seed):
    from fire_b200 import weights as W
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.integers(0, 256, (n // 2, 160, 160, 3), dtype=np.uint8), W.calibration_images(n - n // 2, seed=seed + 50)])


@pytest.fixture(scope="module")
def nets(fire_lib):
    from fire_b200 import engine, weights as W
    out = {}
    for D in (128, 512):
        t = W.synthetic_weights(D, 1234)
        out[D] = (t, engine.FaceNetEngine(D, t))
    return out


def test_config1_facenet128_top1_and_threshold_decisions(nets, oracle_native):
    """BASELINE configs[0]: FaceNet128 on 32 crops + cosine top-1 vs a 10k gallery; ids, 1-d and accept/reject equal."""
    import torch
    from fire_b200.engine import KnnIndex
    from oracle.facenet_ref import facenet_forward, l2_normalize_rows
    tensors, eng = nets[128]
    u8 = _images(32, 0)
    x = u8.astype(np.float32) / 255.0
    raw, l2 = eng.encode_unit_f32(torch.from_numpy(x).cuda())
    ref = facenet_forward(tensors, x)
    got = raw.cpu().numpy()
    assert _cos(got, ref).min() >= 0.9999
    refn = l2_normalize_rows(ref)
    rng = np.random.default_rng(1)
    gal = rng.standard_normal((10_000, 128), dtype=np.float32)
    gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    P = 16                                                    # queries 16..23 are structured images with distinct embeddings
    for j, c in enumerate([0.9, 0.9, 0.9, 0.71, 0.71, 0.69, 0.69, 0.5]):     # plant near-duplicates around thr 0.7
        e = refn[P + j]
        r = rng.standard_normal(128).astype(np.float32); r -= r.dot(e) * e; r /= np.linalg.norm(r)
        gal[100 + j] = c * e + np.sqrt(1 - c * c) * r
    idx = KnnIndex(128, 100000); idx.add(gal)
    dist, ids = idx.search(l2.contiguous(), 1)
    dist, ids = dist.cpu().numpy(), ids.cpu().numpy()
    ora = oracle_native.BFIndexOracle(128); ora.add_items(gal)
    # (a) the matcher alone, on identical inputs (the GPU embeddings): ids bit-exact, distances to fp32 rounding
    sl, sd = ora.knn_query(l2.cpu().numpy(), 1)
    assert np.array_equal(ids, sl.astype(np.int64)) and np.abs(dist - sd).max() < 5e-6
    # (b) the whole chain against the all-CPU chain (fp32 oracle embeddings -> BFIndex): equal wherever the fp16
    #     embedding error (cos >= 0.9999, i.e. up to ~5e-3 in a cosine) cannot matter
    ol, od = ora.knn_query(refn, 2)
    clear = (od[:, 1] - od[:, 0]) > 1e-2
    assert clear.sum() >= 8 and np.array_equal(ids[clear, 0], ol[clear, 0].astype(np.int64))
    assert np.abs(dist[clear, 0] - od[clear, 0]).max() < 5e-3
    thr = 0.7
    accept_gpu, accept_ref = (1 - dist[:, 0]) > thr, (1 - od[:, 0]) > thr    # face_recognition.py:462-463 strict >
    decided = np.abs((1 - od[:, 0]) - thr) > 5e-3
    assert decided[P:P + 8].all() and np.array_equal(accept_gpu[decided], accept_ref[decided])
    assert list(ol[P:P + 8, 0]) == list(range(100, 108))      # every planted row is its query's nearest neighbour
    assert accept_ref[P:P + 5].all() and not accept_ref[P + 5:P + 8].any()   # 0.9/0.71 accepted, 0.69/0.5 rejected


def test_config2_facenet512_batch256(nets):
    """BASELINE configs[1]: FaceNet512, batch 256; every embedding within cosine 0.9999 of the fp32 oracle."""
    import torch
    from oracle.facenet_ref import facenet_forward
    tensors, eng = nets[512]
    u8 = _images(256, 2)
    x = u8.astype(np.float32) / 255.0
    raw, l2 = eng.encode_unit_f32(torch.from_numpy(x).cuda())
    got, gl2 = raw.cpu().numpy(), l2.cpu().numpy()
    sel = np.r_[0:24, 128:152, 232:256]                        # 72 of the 256 through the CPU oracle (seconds, not minutes)
    ref = facenet_forward(tensors, x[sel])
    cos = _cos(got[sel], ref)
    assert cos.min() >= 0.9999, cos.min()
    assert np.abs(np.linalg.norm(gl2, axis=1) - 1).max() < 1e-5
    np.testing.assert_allclose(gl2, got / np.linalg.norm(got, axis=1, keepdims=True), rtol=1e-5, atol=1e-7)
    # batch invariance: the same image gives bit-identical embeddings at B=1, B=3 and B=256
    r1, _ = eng.encode_unit_f32(torch.from_numpy(x[130:131]).cuda())
    r3, _ = eng.encode_unit_f32(torch.from_numpy(x[129:132]).cuda())
    assert np.array_equal(r1.cpu().numpy()[0], got[130]) and np.array_equal(r3.cpu().numpy()[1], got[130])
    # ragged batch sizes (partial tiles everywhere, fewer tiles than SMs / more images than the bench batch)
    for lo, hi in ((100, 137), (0, 211)):
        rb, _ = eng.encode_unit_f32(torch.from_numpy(x[lo:hi]).cuda())
        assert np.array_equal(rb.cpu().numpy(), got[lo:hi])
    big = np.concatenate([x, x[:44]])                           # B = 300
    rbig, _ = eng.encode_unit_f32(torch.from_numpy(big).cuda())
    assert np.array_equal(rbig.cpu().numpy()[:256], got) and np.array_equal(rbig.cpu().numpy()[256:], got[:44])


def test_golden_embeddings(nets):
    """Committed fixture (tests/golden/facenet_golden.json): the synthetic weights exported as a full-structure ONNX graph and
    run by cv2.dnn.readNetFromONNX (make_facenet_golden.py) - numbers neither the oracle nor the engine produced."""
    import torch
    from fire_b200 import weights as W
    with open(os.path.join(GOLDEN, "facenet_golden.json")) as f:
        gold = json.load(f)
    for D in (128, 512):
        tensors, eng = nets[D]
        u8 = W.calibration_images(4, seed=gold["image_seed"])
        raw, _ = eng.encode_unit_f32(torch.from_numpy(u8.astype(np.float32) / 255.0).cuda())
        got = raw.cpu().numpy()
        want = np.asarray(gold[str(D)]["embeddings"], dtype=np.float32)
        assert gold["executor"].startswith("cv2.dnn")
        assert _cos(got, want).min() >= 0.9999


def test_unfused_plan_matches_fused(fire_lib):
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 99)
    a, b = engine.FaceNetEngine(128, t, fuse_siblings=True), engine.FaceNetEngine(128, t, fuse_siblings=False)
    x = torch.from_numpy(_images(6, 9).astype(np.float32) / 255.0).cuda()
    ra, _ = a.encode_unit_f32(x); rb, _ = b.encode_unit_f32(x)
    # horizontal fusion regroups output channels (bit-identical); the grouped Block35 3x3 launch also stores the branch
    # concat in another channel order, which permutes the K summation order of the up conv: equal to fp32 rounding
    ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
    assert _cos(ra, rb).min() >= 0.99999 and np.abs(ra - rb).max() <= 1e-2 * np.abs(rb).max()


def test_strip_kernel_matches_gather_kernel(fire_lib, monkeypatch):
    """conv_strip_kernel (halo patch + shifted descriptors, flat tiles over pitched buffers) against the im2col-gather
    path of conv_igemm_kernel on the same weights: same products, same fp32 accumulation, only the order in which the
    bias joins the sum differs."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 7)
    x = torch.from_numpy(_images(5, 11).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(128, t)                                   # strip kernels on (default)
    ra, _ = a.encode_unit_f32(x)
    monkeypatch.setenv("FIRE_B200_STRIP", "0")
    b = engine.FaceNetEngine(128, t, pitched=False)                    # every k x k layer through the gather path
    rb, _ = b.encode_unit_f32(x)
    ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
    assert _cos(ra, rb).min() >= 0.99999          # one-ulp fp16 flips (bias joins the fp32 sum first vs last) through ~100 layers
    assert np.abs(ra - rb).max() <= 1e-2 * np.abs(rb).max()


@pytest.mark.parametrize("B", [1, 5, 64, 301])
def test_block17_fused_chain_matches_layer_by_layer(fire_lib, monkeypatch, B):
    """block17_fused_kernel (the ten Block17 blocks in one launch: intermediates in shared memory, 1x7 / 7x1 as
    row-shifted windows, residual added in the epilogue) against the same plan run layer by layer
    (FIRE_B200_FUSE17=0: conv_igemm_kernel x 40).  Same fp16 operands, same fp32 accumulation; bias and residual join
    the sum in a different order.  B = 1 / 5 / 301: a half-empty tile (TMA zero fill and clipping); 301: CTAs that own two tiles."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 5)
    x = torch.from_numpy(_images(B, 13).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(128, t)
    ra, _ = a.encode_unit_f32(x)
    ra2, _ = a.encode_unit_f32(x)
    assert torch.equal(ra, ra2)                                    # deterministic
    monkeypatch.setenv("FIRE_B200_FUSE17", "0")
    b = engine.FaceNetEngine(128, t)
    assert b.num_launches - a.num_launches == 39                   # 40 conv launches became one
    rb, _ = b.encode_unit_f32(x)
    ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
    assert np.isfinite(ra).all()
    assert _cos(ra, rb).min() >= 0.99999
    assert np.abs(ra - rb).max() <= 1e-2 * np.abs(rb).max()


@pytest.mark.parametrize("B", [1, 7, 149, 300])
def test_block35_fused_chain_matches_layer_by_layer(fire_lib, monkeypatch, B):
    """block35_fused_kernel (the five Block35 blocks in one launch, one image per CTA pass: 3x3 convs as row-shifted
    windows of a pitched zero-bordered copy in shared memory) against the same plan run layer by layer
    (FIRE_B200_FUSE35=0: conv_igemm_kernel x 20).  B = 149 / 300: CTAs that own two / three images."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 6)
    x = torch.from_numpy(_images(B, 17).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(128, t)
    ra, _ = a.encode_unit_f32(x)
    ra2, _ = a.encode_unit_f32(x)
    assert torch.equal(ra, ra2)
    if B > 148:                                                    # the balanced task order (chains wander over CTAs, per-(block, image)
        monkeypatch.setenv("FIRE_B200_B35_BALANCE", "0")           # flags) against "a CTA owns whole images": same arithmetic
        c = engine.FaceNetEngine(128, t)
        monkeypatch.delenv("FIRE_B200_B35_BALANCE")
        rc, _ = c.encode_unit_f32(x)
        assert torch.equal(ra, rc)
        c.close()
    monkeypatch.setenv("FIRE_B200_FUSE35", "0")
    b = engine.FaceNetEngine(128, t)
    assert b.num_launches - a.num_launches == 19                   # 20 conv launches became one
    rb, _ = b.encode_unit_f32(x)
    ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
    assert np.isfinite(ra).all()
    assert _cos(ra, rb).min() >= 0.99999
    assert np.abs(ra - rb).max() <= 1e-2 * np.abs(rb).max()


@pytest.mark.parametrize("B", [1, 3, 130])
def test_pair_stem_matches_pixel_stem(fire_lib, B):
    """Conv2d_1a / Conv2d_2a run over PIXEL PAIRS (netplan.Plan.pair_stem: N = 64 instead of 32, half as many tiles, the
    weights shifted copies of the same taps) against the same strip kernel over single pixels: the stored activations of
    both layers - every valid pixel, including column 78 of Conv2d_1a, which the pair conv produces in the flat tiles'
    spare column - and the embeddings.  Same fp16 operands; the extra zero-weight taps add exact zeros."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(512, 21)
    x = torch.from_numpy(_images(B, 23).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(512, t, reuse_buffers=False)
    b = engine.FaceNetEngine(512, t, reuse_buffers=False, pair_stem=False)
    assert a.plan.pair_stem and not b.plan.pair_stem
    xa, xb = a.ingest_unit_f32(x), b.ingest_unit_f32(x)
    ra, _ = a.forward(xa)
    rb, _ = b.forward(xb)
    for i in (1, 2):                                                   # outputs of Conv2d_1a, Conv2d_2a
        ba, bb = a.read_buffer(a.plan.ops[i - 1].dst.buf, xa), b.read_buffer(b.plan.ops[i - 1].dst.buf, xb)
        assert ba.shape == bb.shape and ba.shape[2] in (79, 77)
        assert np.abs(ba - bb).max() <= 2e-3 * max(1.0, np.abs(bb).max()), (i, np.abs(ba - bb).max())
    ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
    assert _cos(ra, rb).min() >= 0.99999
    a.close(); b.close()


@pytest.mark.parametrize("B", [1, 12, 150])
def test_pool_conv_fused_matches_two_launches(fire_lib, monkeypatch, B):
    """pool_conv_fused_kernel (MaxPool_3a + Conv2d_3b in one launch: the TMA brings seven input rows per tile, eight warps
    reduce the 3 x 3 / 2 windows from shared memory into the MMA operand, so the pooled tensor never exists in memory) against
    maxpool3x3s2_kernel + conv_igemm_kernel (FIRE_B200_FUSE_POOL=0): the stored output of Conv2d_3b - every position, including
    the two-row last tile of an image - and the embeddings.  The max is exact; the bias joins the sum in fp32 here and as an fp16
    hi + lo MMA there."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(512, 27)
    x = torch.from_numpy(_images(B, 29).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(512, t, reuse_buffers=False)
    monkeypatch.setenv("FIRE_B200_FUSE_POOL", "0")
    b = engine.FaceNetEngine(512, t, reuse_buffers=False)
    monkeypatch.delenv("FIRE_B200_FUSE_POOL")
    assert b.num_launches - a.num_launches == 1
    xa, xb = a.ingest_unit_f32(x), b.ingest_unit_f32(x)
    ra, _ = a.forward(xa)
    rb, _ = b.forward(xb)
    i3b = next(i for i, o in enumerate(a.plan.ops) if o.label == "Conv2d_3b_1x1")
    ba, bb = a.read_buffer(a.plan.ops[i3b].dst.buf, xa), b.read_buffer(b.plan.ops[i3b].dst.buf, xb)
    assert ba.shape == bb.shape == (B, 38, 38, 80)
    assert np.abs(ba - bb).max() <= 2e-3 * max(1.0, np.abs(bb).max()), np.abs(ba - bb).max()
    ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
    assert np.isfinite(ra).all() and _cos(ra, rb).min() >= 0.99999
    a.close(); b.close()


@pytest.mark.parametrize("B", [1, 13, 14, 100, 256, 300])
def test_block8_fused_tail_matches_layer_by_layer(fire_lib, monkeypatch, B):
    """block8_fused_kernel (1x3 -> 3x1 -> up + residual of a Block8 block in one launch: a CTA owns 13 images x 256 output
    channels, the 1x3 taps are TMA boxes shifted in x with out-of-bounds zero fill, the 3x1 taps row-shifted windows of a
    y-major buffer, the residual one more K block) against the same plan run layer by layer (FIRE_B200_FUSE8=0:
    conv_igemm_kernel x 3 per block).  Same fp16 operands and fp32 accumulation; the bias joins the sum in fp32 instead of
    as an fp16 hi+lo MMA.  B = 1 / 14 / 100: ragged last image group (TMA zero fill and clipping); 300: two waves of CTAs."""
    import torch
    from fire_b200 import engine, weights as W
    for D in (128, 512):
        t = W.synthetic_weights(D, 8)
        x = torch.from_numpy(_images(B, 19).astype(np.float32) / 255.0).cuda()
        a = engine.FaceNetEngine(D, t)
        ra, _ = a.encode_unit_f32(x)
        ra2, _ = a.encode_unit_f32(x)
        assert torch.equal(ra, ra2)                                    # deterministic
        monkeypatch.setenv("FIRE_B200_FUSE8", "0")
        b = engine.FaceNetEngine(D, t)
        monkeypatch.delenv("FIRE_B200_FUSE8")
        assert b.num_launches - a.num_launches == 13                   # 3 launches per block became one, and the last one also does the average pool
        rb, _ = b.encode_unit_f32(x)
        ra, rb = ra.cpu().numpy(), rb.cpu().numpy()
        assert np.isfinite(ra).all()
        assert _cos(ra, rb).min() >= 0.99999
        assert np.abs(ra - rb).max() <= 1e-2 * np.abs(rb).max()
        a.close(); b.close()


@pytest.mark.parametrize("B", [149, 256, 300])
def test_repeated_forwards_are_bit_identical(fire_lib, B):
    """Stress for the cross-CTA hand-overs (block35_fused's wandering chains publish y through release / acquire flags; the
    flags carry the launch number and are never reset): 60 back-to-back forwards, alternating with a second batch so every
    buffer is overwritten in between, must reproduce the first result bit for bit."""
    import torch
    from fire_b200 import engine, weights as W
    eng = engine.FaceNetEngine(512, W.synthetic_weights(512, 3, calibrate=False))
    xa = eng.ingest_unit_f32(torch.from_numpy(_images(B, 41).astype(np.float32) / 255.0).cuda())
    xb = eng.ingest_unit_f32(torch.from_numpy(_images(B, 42).astype(np.float32) / 255.0).cuda())
    ra, _ = eng.forward(xa)
    rb, _ = eng.forward(xb)
    ra, rb = ra.clone(), rb.clone()
    assert torch.isfinite(ra).all() and not torch.equal(ra, rb)
    for i in range(60):
        r, _ = eng.forward(xa if i % 2 == 0 else xb)
        assert torch.equal(r, ra if i % 2 == 0 else rb), i
    eng.close()


@pytest.mark.parametrize("B", [2, 37, 130])
def test_im2col_tma_operand_is_bit_identical_to_gather(fire_lib, monkeypatch, B):
    """k x k layers with Cin % 64 == 0 (Conv2d_4b, Mixed_6a, Mixed_7a) fetch their A operand with im2col-mode TMA loads
    (one instruction per (tap, 64-channel slice) and 128 output pixels; padding, stride and the ragged last tile are the
    tensor map's business) instead of the cp.async gather warps: the same bytes land in the same swizzled tile, so the
    embeddings must be bit-identical."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 10)
    x = torch.from_numpy(_images(B, 31).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(128, t)
    ra, _ = a.encode_unit_f32(x)
    monkeypatch.setenv("FIRE_B200_IM2COL", "0")
    b = engine.FaceNetEngine(128, t)
    rb, _ = b.encode_unit_f32(x)
    assert torch.equal(ra, rb)


@pytest.mark.parametrize("B", [2, 9, 150])
def test_strip_kernel_cta_pairs_are_bit_identical(fire_lib, monkeypatch, B):
    """conv_strip_kernel_t<true> (FIRE_B200_STRIP_PAIR=1): clusters of two CTAs take the same position block of two
    consecutive images; ONE M = 256 tcgen05.mma.cta_group::2 of the leader covers both tiles, each CTA holds half of the
    weight rows, completions are multicast to both CTAs and the peer reports through remote mbarrier arrives.  Same
    products in the same order per output row: bit-identical (odd B: the last image's partner is a zero-filled phantom)."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 12)
    x = torch.from_numpy(_images(B, 37).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(128, t)
    ra, _ = a.encode_unit_f32(x)
    monkeypatch.setenv("FIRE_B200_STRIP_PAIR", "1")
    b = engine.FaceNetEngine(128, t)
    rb, _ = b.encode_unit_f32(x)
    assert torch.equal(ra, rb)


@pytest.mark.parametrize("B", [2, 65, 256])
def test_igemm_cta_pairs_are_bit_identical(fire_lib, monkeypatch, B):
    """conv_igemm_kernel_t<true>: the TMA-fed layers without residual (Conv2d_3b / 4b, Mixed_6a, Mixed_7a at large batch) run
    as CTA pairs - M = 256 per tcgen05.mma.cta_group::2, each CTA loads its own activation tile and half of the weight
    rows.  Per output row the same products in the same order as the single-CTA kernel (FIRE_B200_IGEMM_PAIR=0)."""
    import torch
    from fire_b200 import engine, weights as W
    t = W.synthetic_weights(128, 14)
    x = torch.from_numpy(_images(B, 41).astype(np.float32) / 255.0).cuda()
    a = engine.FaceNetEngine(128, t)
    ra, _ = a.encode_unit_f32(x)
    monkeypatch.setenv("FIRE_B200_IGEMM_PAIR", "0")
    b = engine.FaceNetEngine(128, t)
    rb, _ = b.encode_unit_f32(x)
    assert torch.equal(ra, rb)


def test_crop_encode_pipeline_equals_direct_path(nets):
    """The streaming public call (pinned host crops -> H2D -> K1 -> K2 -> D2H, double-buffered) returns exactly what
    the step-by-step path returns, for every in-flight batch."""
    import torch
    from fire_b200 import _lib, engine
    _, eng = nets[512]
    B = 8
    batches = [torch.from_numpy(_images(B, 20 + i)).pin_memory() for i in range(5)]
    pipe = engine.CropEncodePipeline(eng, B, depth=2, normalize=True)
    got = []
    for i, b in enumerate(batches):
        t = pipe.submit(b)
        if i >= 1:
            got.append(pipe.result(t - 1).clone())
    got.append(pipe.result(len(batches) - 1).clone())
    boxes = torch.tensor([[0, 0, 160, 160]] * B, dtype=torch.int32, device="cuda")
    fid = torch.arange(B, dtype=torch.int32, device="cuda")
    desc = torch.tensor([[i * 76800, 160, 160, 480] for i in range(B)], dtype=torch.int64, device="cuda")
    for b, g in zip(batches, got):
        f16, _, _ = engine.preprocess_boxes(b.cuda(), desc, boxes, fid, _lib.PRE_REFERENCE, True, False)
        _, l2 = eng.forward(f16, want_l2=True)
        assert torch.equal(l2.cpu(), g)


def test_forward_argument_errors(nets):
    import torch
    from fire_b200 import _lib
    from fire_b200._lib import FireError
    _, eng = nets[128]
    x = torch.zeros(1, 80, 80, 16, dtype=torch.float16, device="cuda")
    out = torch.zeros(1, 128, device="cuda")
    with pytest.raises(FireError):
        _lib.check(_lib.lib().fire_facenet_forward(eng._h, x.data_ptr(), 1, out.data_ptr(), None, x.data_ptr(), 16, None))


def test_no_activation_saturates_fp16(fire_lib, monkeypatch):
    """fp16 storage has a range (65504) that bf16 does not: the epilogues convert with .satfinite, which would clip SILENTLY.
    Saturation counter: with every activation kept in its own memory (reuse_buffers=False) and the residual chains run layer
    by layer (so their branch tensors exist), every stored activation of a forward is read back and counted against the
    limit.  The count must be 0 and the headroom is reported (profiles/r02_bf16_vs_fp16.txt has the CPU-side table)."""
    import torch
    from fire_b200 import engine, weights as W
    monkeypatch.setenv("FIRE_B200_FUSE17", "0")
    monkeypatch.setenv("FIRE_B200_FUSE35", "0")
    monkeypatch.setenv("FIRE_B200_FUSE8", "0")
    monkeypatch.setenv("FIRE_B200_FUSE_POOL", "0")
    for D in (512, 128):
        t = W.synthetic_weights(D, 1234)
        eng = engine.FaceNetEngine(D, t, reuse_buffers=False)
        u8 = np.concatenate([_images(6, 31), np.full((1, 160, 160, 3), 255, np.uint8), np.zeros((1, 160, 160, 3), np.uint8)])   # incl. all-white / all-black
        x = eng.ingest_unit_f32(torch.from_numpy(u8.astype(np.float32) / 255.0).cuda())
        raw, _ = eng.forward(x)
        assert torch.isfinite(raw).all()
        worst, n_sat, n_read = 0.0, 0, 0
        for i, b in enumerate(eng.plan.bufs):
            if b.external or b.elt != 2:
                continue
            a = eng.read_buffer(i, x)
            assert np.isfinite(a).all(), i
            worst = max(worst, float(np.abs(a).max()))
            n_sat += int((np.abs(a) >= 65504.0).sum())
            n_read += 1
        assert n_read >= 60 and n_sat == 0, (n_read, n_sat)
        assert worst < 65504.0 / 16, worst                           # more than 4 bits of headroom on this workload
        eng.close()
