"""Aligned-crop path: the C oracle of cv2.warpAffine and BOTH getAffineTransform restatements
(oracle C, product Python) are pinned bit for bit against the live cv2 of this image."""
import cv2
import numpy as np


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    for t in range(n):
        H, W = int(rng.integers(120, 700)), int(rng.integers(160, 900))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if t % 3:
            img = cv2.GaussianBlur(img, (0, 0), 1.5)
        cx, cy = rng.uniform(-40, W + 40), rng.uniform(-40, H + 40)
        s, ang = rng.uniform(15, 200), rng.uniform(-0.6, 0.6)
        pts = np.float32([(cx - s * np.cos(ang), cy - s * np.sin(ang)), (cx + s * np.cos(ang), cy + s * np.sin(ang)),
                          (cx + 0.7 * s * np.sin(ang) + rng.uniform(-5, 5), cy + 0.7 * s * np.cos(ang) + rng.uniform(-5, 5))])
        if t % 4 == 0:
            pts = np.round(pts)                       # the detectors emit int32 landmarks (yunet_face_detector.py:53)
        yield img, pts


def test_get_affine_transform_matches_cv2_bitwise(oracle_native):
    from fire_b200.preprocess import ALIGN_DST, get_affine_transform
    for _, pts in _cases(200, 0):
        want = cv2.getAffineTransform(pts, ALIGN_DST)
        assert np.array_equal(oracle_native.get_affine_transform(pts, ALIGN_DST), want)
        assert np.array_equal(get_affine_transform(pts, ALIGN_DST), want)          # the product's host-side solver
    assert np.array_equal(ALIGN_DST, np.float32([(56, 56), (104, 56), (80, 88)]))


def test_warp_affine_oracle_matches_cv2_bitwise(oracle_native):
    from fire_b200.preprocess import ALIGN_DST
    for img, pts in _cases(60, 1):
        M = cv2.getAffineTransform(pts, ALIGN_DST)
        assert np.array_equal(oracle_native.warp_affine(img, M), cv2.warpAffine(img, M, (160, 160)))


def test_degenerate_landmarks_give_zero_matrix(oracle_native):
    from fire_b200.preprocess import ALIGN_DST, get_affine_transform
    pts = np.float32([(10, 10), (20, 20), (30, 30)])                                # collinear
    assert np.array_equal(get_affine_transform(pts, ALIGN_DST), cv2.getAffineTransform(pts, ALIGN_DST))
