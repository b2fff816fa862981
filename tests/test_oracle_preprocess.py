"""Pins the preprocessing oracle (oracle/fire_oracle.c) bit-for-bit against the live cv2 in this image, i.e. against
the very call the reference makes (modules/encoder.py:20), and checks the crop rule of face_recognition.py:412-420."""
import cv2
import numpy as np
import pytest

SHAPES = [(160, 160), (320, 320), (480, 320), (320, 480), (640, 480), (233, 201), (161, 160), (200, 399), (97, 83),
          (120, 159), (52, 47), (400, 100), (100, 400), (48, 48), (161, 500), (1080, 1920), (159, 161), (300, 160),
          (160, 300), (333, 333), (80, 80), (40, 640), (480, 480), (800, 800), (960, 640), (1, 1), (2, 500), (159, 159)]


@pytest.mark.parametrize("shape", SHAPES)
def test_resize_area_equals_cv2(oracle_native, shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    if shape[0] > 8 and shape[1] > 8 and (shape[0] + shape[1]) % 2:
        img = cv2.GaussianBlur(img, (0, 0), 3)
    want = cv2.resize(img, (160, 160), interpolation=cv2.INTER_AREA)
    assert np.array_equal(oracle_native.resize_area(img), want)


def test_resize_area_random_shapes(oracle_native):
    rng = np.random.default_rng(0)
    for _ in range(60):
        h, w = int(rng.integers(20, 700)), int(rng.integers(20, 700))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(oracle_native.resize_area(img), cv2.resize(img, (160, 160), interpolation=cv2.INTER_AREA)), (h, w)


def test_strided_view_like_a_frame_crop(oracle_native):
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    view = frame[20:241, 30:350]                                   # non-contiguous rows, as image[y:y+h, x:x+w] is
    assert np.array_equal(oracle_native.resize_area(view), cv2.resize(view, (160, 160), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("box", [[10, 20, 100, 120], [-15, -30, 200, 180], [350, 250, 200, 200], [0, 0, 400, 300],
                                 [390, 290, 50, 50], [500, 10, 40, 40], [10, 10, 0, 50], [10, 10, 50, -3], [-50, -50, 40, 40]])
def test_crop_rule_matches_reference_slicing(oracle_native, box):
    """x,y,w,h are clamped to >= 0 independently (a negative x is zeroed, w is NOT shrunk), numpy clips the far edge."""
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    x, y, w, h = box
    x, y, w, h = max(0, x), max(0, y), max(0, w), max(0, h)       # face_recognition.py:413-416
    face = frame[y:y + h, x:x + w]                                 # :417
    rc, u8, f32 = oracle_native.crop_preprocess(frame, box)
    if face.size == 0:                                             # :418-420 -> the face is skipped
        assert rc == 1
        return
    assert rc == 0
    resized = cv2.resize(face, (160, 160), interpolation=cv2.INTER_AREA)          # encoder.py:20
    img = resized.astype(np.float32) / 255.0                                      # encoder.py:21
    assert np.array_equal(u8, resized) and np.array_equal(f32, img)
