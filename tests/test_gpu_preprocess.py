"""GPU parity: fused crop/resize/normalise kernel (fire_preprocess) vs the cv2-exact oracle and live cv2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(h, w, seed):
    import cv2
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return cv2.GaussianBlur(img, (0, 0), 1.5) if seed % 2 else img


def test_reference_mode_bit_exact(fire_lib, oracle_native):
    import cv2
    import torch
    from fire_b200 import _lib, engine
    frames = [_frame(1080, 1920, 1), _frame(480, 640, 2), _frame(200, 300, 3)]
    rng = np.random.default_rng(0)
    boxes, bf = [], []
    fixed = [(0, [100, 100, 160, 160]), (0, [200, 50, 320, 320]), (0, [500, 300, 233, 201]), (0, [10, 10, 97, 83]),
             (0, [1800, 900, 400, 400]), (0, [-20, -30, 200, 180]), (1, [0, 0, 640, 480]), (1, [300, 200, 480, 480]),
             (2, [50, 50, 52, 47]), (2, [0, 0, 0, 10]), (2, [400, 10, 50, 50]), (1, [-100, 5, 80, 90]), (0, [64, 64, 480, 960]),
             (0, [0, 0, 1920, 1080]), (2, [10, 10, 40, 160]), (1, [5, 5, 161, 159])]
    for f, b in fixed:
        boxes.append(b); bf.append(f)
    for _ in range(48):                                        # YuNet-style random boxes, some crossing the frame edge
        f = int(rng.integers(0, 3)); H, W = frames[f].shape[:2]
        boxes.append([int(rng.integers(-40, W)), int(rng.integers(-40, H)), int(rng.integers(20, 420)), int(rng.integers(20, 420))])
        bf.append(f)
    flat, desc = engine.frames_to_device(frames)
    bt = torch.tensor(boxes, dtype=torch.int32).cuda(); ft = torch.tensor(bf, dtype=torch.int32).cuda()
    f16, f32, status = engine.preprocess_boxes(flat, desc, bt, ft, _lib.PRE_REFERENCE, True, True)
    f16, f16_pad = engine.network_input_to_pixels(f16)          # space-to-depth network input -> [n,160,160,3] pixels + padding
    f16 = f16.cpu().numpy(); f16_pad = f16_pad.cpu().numpy(); f32 = f32.cpu().numpy(); status = status.cpu().numpy()
    n_live = 0
    for i, (b, f) in enumerate(zip(boxes, bf)):
        rc, u8, want = oracle_native.crop_preprocess(frames[f], b)
        assert status[i] == rc
        if rc:
            assert not f32[i].any() and not f16[i].any()
            continue
        assert np.array_equal(f32[i], want), (i, b)                                  # bit-exact float32 (u8 / 255)
        assert np.array_equal(f16[i], u8.astype(np.float32)) and not f16_pad[i].any()
        x, y, w, h = (max(0, v) for v in b)                                          # face_recognition.py:412-417
        crop = frames[f][y:y + h, x:x + w]
        live = cv2.resize(crop, (160, 160), interpolation=cv2.INTER_AREA).astype(np.float32) / 255.0
        assert np.array_equal(f32[i], live)
        n_live += 1
    assert n_live > 40


def test_swap_rb_and_northstar(fire_lib):
    import torch
    from fire_b200 import _lib, engine
    frame = _frame(400, 500, 5)
    boxes = [[20, 30, 300, 250], [100, 100, 90, 120], [0, 0, 500, 400]]
    flat, desc = engine.frames_to_device([frame])
    bt = torch.tensor(boxes, dtype=torch.int32).cuda(); ft = torch.zeros(3, dtype=torch.int32).cuda()
    _, a, _ = engine.preprocess_boxes(flat, desc, bt, ft, _lib.PRE_REFERENCE, False, True)
    _, b, _ = engine.preprocess_boxes(flat, desc, bt, ft, _lib.PRE_REFERENCE | _lib.PRE_FLAG_SWAP_RB, False, True)
    assert torch.equal(a.flip(-1), b)
    # north-star mode: float half-pixel bilinear + prewhiten, against a numpy restatement (tolerance: fp32 order)
    f16, y, _ = engine.preprocess_boxes(flat, desc, bt, ft, _lib.PRE_NORTHSTAR, True, True)
    y = y.cpu().numpy(); f16 = engine.network_input_to_pixels(f16)[0].cpu().numpy()
    for i, (x0, y0, w, h) in enumerate(boxes):
        crop = frame[y0:y0 + h, x0:x0 + w].astype(np.float32)
        fx = np.clip((np.arange(160, dtype=np.float32) + 0.5) * np.float32(w / 160) - 0.5, 0, w - 1)
        fy = np.clip((np.arange(160, dtype=np.float32) + 0.5) * np.float32(h / 160) - 0.5, 0, h - 1)
        xa, ya = fx.astype(int), fy.astype(int)
        xb, yb = np.minimum(xa + 1, w - 1), np.minimum(ya + 1, h - 1)
        tx, ty = (fx - xa)[None, :, None], (fy - ya)[:, None, None]
        top = (1 - tx) * crop[ya][:, xa] + tx * crop[ya][:, xb]
        bot = (1 - tx) * crop[yb][:, xa] + tx * crop[yb][:, xb]
        img = (1 - ty) * top + ty * bot
        mean, std = img.astype(np.float64).mean(), img.astype(np.float64).std()
        want = (img - mean) / max(std, 1.0 / np.sqrt(img.size))
        assert np.abs(y[i] - want).max() < 2e-4                                      # tolerance: fp32 vs fp64 statistics
        assert abs(float(y[i].mean())) < 1e-4 and abs(float(y[i].std()) - 1.0) < 1e-3
        assert np.abs(f16[i] - 255.0 * want).max() < 0.5 + 255.0 * 2e-4 + 1.0   # fp16 rounding of 255*y (|.|<2048 -> ulp<=1)


def test_aligned_crops_match_cv2_warp_affine_bitwise(fire_lib):
    """SURVEY 8(f) row 1 - the enrol path's aligned crop (yunet_face_detector.py:135-165): getAffineTransform +
    warpAffine(image, M, (160,160)) + [:, :, ::-1], on the B200, bit for bit against the live cv2."""
    import cv2
    import torch
    from fire_b200 import engine
    from fire_b200.preprocess import ALIGN_DST, align_faces
    rng = np.random.default_rng(9)
    frames = [_frame(1080, 1920, 1), _frame(480, 640, 2), _frame(200, 300, 3)]
    lms, ff = [], []
    for t in range(40):
        f = t % 3
        H, W = frames[f].shape[:2]
        cx, cy = rng.uniform(-30, W + 30), rng.uniform(-30, H + 30)           # some faces hang over the frame edge
        s, ang = rng.uniform(12, 0.45 * min(H, W)), rng.uniform(-0.7, 0.7)
        pts = np.float32([(cx - s * np.cos(ang), cy - s * np.sin(ang)), (cx + s * np.cos(ang), cy + s * np.sin(ang)),
                          (cx + 0.7 * s * np.sin(ang), cy + 0.7 * s * np.cos(ang))])
        lms.append(np.round(pts) if t % 2 else pts)
        ff.append(f)
    got = align_faces(frames, np.stack(lms), ff, swap_rb=True, output="uint8")
    for i in range(len(lms)):
        M = cv2.getAffineTransform(lms[i], ALIGN_DST)
        want = cv2.warpAffine(frames[ff[i]], M, (160, 160))[:, :, ::-1]
        assert np.array_equal(got[i], want), i
    # the network input of those crops = preprocess_for_encoder(crop) * 255 (a 160x160 crop is copied by INTER_AREA)
    f16 = align_faces(frames, np.stack(lms), ff, swap_rb=True, output="device")
    pix, pad = engine.network_input_to_pixels(f16)
    assert np.array_equal(pix.cpu().numpy(), got.astype(np.float32)) and not pad.any()
    assert not torch.isnan(f16).any()


def test_staged_loads_unaligned_frames_and_roi_upload(fire_lib, oracle_native):
    """The kernel stages source rows with 16-byte loads: frames that start at an odd byte (chunks hanging over the first /
    last byte of the allocation) still give the oracle's bytes; and the ROI upload (fire_pack_rois_host + ONE copy) is
    bit-identical to uploading whole frames (configs[4])."""
    import torch
    from fire_b200 import _lib, engine
    rng = np.random.default_rng(21)
    frames = [_frame(270, 480, 4), _frame(128, 208, 7)]                 # strides 1440 and 624: multiples of 16 -> staged path
    boxes = np.array([[0, 0, 160, 160], [0, 0, 480, 270], [320, 110, 160, 160], [300, 100, 200, 200], [-10, -10, 100, 90], [470, 260, 50, 50],
                      [0, 0, 208, 128], [48, 0, 160, 128], [100, 64, 108, 64], [5, 5, 33, 41], [600, 5, 20, 20]] +
                     [[int(rng.integers(-20, 440)), int(rng.integers(-20, 250)), int(rng.integers(20, 300)), int(rng.integers(20, 300))] for _ in range(30)],
                     dtype=np.int32)
    bf = np.array([0] * 6 + [1] * 5 + [0] * 30, dtype=np.int32)
    prefix = 5                                                           # frames start at byte 5 of the allocation
    flat = np.concatenate([np.zeros(prefix, np.uint8)] + [f.reshape(-1) for f in frames])
    desc = np.array([[prefix, 270, 480, 1440], [prefix + frames[0].size, 128, 208, 624]], dtype=np.int64)
    dev = torch.from_numpy(flat).cuda()
    bt, ft = torch.from_numpy(boxes).cuda(), torch.from_numpy(bf).cuda()
    f16, f32, status = engine.preprocess_boxes(dev, torch.from_numpy(desc).cuda(), bt, ft, _lib.PRE_REFERENCE, True, True)
    torch.cuda.synchronize()
    for i, b in enumerate(boxes):
        rc, u8, want = oracle_native.crop_preprocess(frames[bf[i]], b)
        assert int(status[i]) == rc and np.array_equal(f32[i].cpu().numpy(), want), (i, b)
    # ROI upload: same boxes through the stager, twice (both slots), bit-identical outputs
    stager = engine.RoiStager(max_bytes=8 << 20, depth=2)
    host_flat = torch.from_numpy(flat).pin_memory()
    for it in range(5):                                                  # with and without the background pack of the next step
        d_frames, d_desc, d_boxes, d_bf = stager.submit(host_flat, desc, boxes, bf, next_args=(host_flat, desc, boxes, bf) if it % 2 == 0 else None)
        g16, g32, gst = engine.preprocess_boxes(d_frames, d_desc, d_boxes, d_bf, _lib.PRE_REFERENCE, True, True)
        stager.release()
        torch.cuda.synchronize()
        assert torch.equal(g16, f16) and torch.equal(g32, f32) and torch.equal(gst, status)
    stager.close()
    dma = engine.RoiStager(max_bytes=8 << 20, depth=2, mode="dma")      # the copy engine gathers: one 2-D copy per rectangle
    for _ in range(3):
        d_frames, d_desc, d_boxes, d_bf = dma.submit(host_flat, desc, boxes, bf)
        g16, g32, gst = engine.preprocess_boxes(d_frames, d_desc, d_boxes, d_bf, _lib.PRE_REFERENCE, True, True)
        dma.release()
        torch.cuda.synchronize()
        assert torch.equal(g16, f16) and torch.equal(g32, f32) and torch.equal(gst, status)
    assert dma.last_bytes == stager.last_bytes
    assert stager.last_bytes < flat.nbytes * 6                           # rectangles, not frames (boxes overlap, so not < 1x here)
