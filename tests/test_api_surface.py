"""The drop-in surfaces keep the reference's names, argument lists and defaults (tests/golden/reference_api.json is
extracted from /root/reference by tests/golden/make_reference_api.py), and their host-side behaviour matches."""
import inspect
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_api.json")))


def _args(fn):
    return [p.name for p in inspect.signature(fn).parameters.values()]


def _check_function(ref, fn, allow_extra=False):
    have = _args(fn)
    if allow_extra:                      # additive keyword arguments after the reference's own are allowed
        assert have[:len(ref["args"])] == ref["args"], (fn, have, ref["args"])
    else:
        assert have == ref["args"], (fn, have, ref["args"])
    params = list(inspect.signature(fn).parameters.values())[:len(ref["args"])]
    defaults = [repr(p.default) for p in params if p.default is not inspect.Parameter.empty]
    assert [d.replace('"', "'") for d in defaults] == [d.replace('"', "'") for d in ref["defaults"]], (fn, defaults, ref["defaults"])


def test_encoder_surface():
    from fire_b200 import encoder
    for name, ref in GOLD["encoder"]["classes"]["Encoder"].items():
        _check_function(ref, getattr(encoder.Encoder, name))


def test_facenet_gpu_surface():
    from fire_b200 import facenet_gpu
    for name, ref in GOLD["facenet_gpu"]["functions"].items():
        _check_function(ref, getattr(facenet_gpu, name))
    for name, ref in GOLD["facenet_gpu"]["classes"]["FaceNetClient"].items():
        _check_function(ref, getattr(facenet_gpu.FaceNetClient, name))
    assert facenet_gpu.scaling(np.array([2.0]), 0.17)[0] == pytest.approx(0.34)
    with pytest.raises(ValueError, match="Invalid mode selected"):
        facenet_gpu.load_facenet_model("weights/facenet128.onnx", mode="tpu")          # facenet_gpu.py:59-60


def test_missing_and_lfs_pointer_weights_raise_like_the_reference(tmp_path, monkeypatch):
    from fire_b200 import facenet_gpu
    monkeypatch.setenv("FIRE_B200_WEIGHTS_DIR", str(tmp_path))
    monkeypatch.delenv("FIRE_B200_SYNTHETIC_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError, match="ONNX model not found"):               # facenet_gpu.py:36-37
        facenet_gpu.load_facenet_model("weights/facenet512.onnx", mode="cpu_optimized")
    os.makedirs(tmp_path / "weights")
    (tmp_path / "weights" / "facenet512.onnx").write_text(
        "version https://git-lfs.github.com/spec/v1\noid sha256:f0dfb218\nsize 94037431\n")
    with pytest.raises(ValueError, match="not corrupted"):                              # facenet_gpu.py:75-79
        facenet_gpu.load_facenet_model("weights/facenet512.onnx", mode="cpu_optimized")


def test_hnsw_manager_surface():
    from fire_b200 import hnsw_manager
    for name, ref in GOLD["hnsw_manager"]["classes"]["HNSWManager"].items():
        _check_function(ref, getattr(hnsw_manager.HNSWManager, name), allow_extra=(name == "__init__"))


def test_preprocess_surface_and_host_functions(tmp_path):
    import cv2
    from fire_b200 import preprocess
    for name, ref in GOLD["preprocess"]["functions"].items():
        _check_function(ref, getattr(preprocess, name))
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (300, 500, 3), dtype=np.uint8)
    # resize_image / preprocess_image: same numbers as processing/preprocess.py:86-145 (restated inline)
    out, scale = preprocess.resize_image(img, [1024, 1980], True)
    assert scale == 1024 / 300.0 and out.shape[:2] == cv2.resize(img, None, None, fx=scale, fy=scale).shape[:2]
    out, scale = preprocess.resize_image(img, [1024, 1980], False)
    assert scale == 1.0 and out is img
    wide = rng.integers(0, 256, (100, 900, 3), dtype=np.uint8)
    _, scale = preprocess.resize_image(wide, [1024, 1980], True)
    assert scale == 1980 / 900.0                                                       # long side capped at 1980
    t, hw, s = preprocess.preprocess_image(img, False)
    assert t.shape == (1, 300, 500, 3) and t.dtype == np.float32 and hw == (300, 500) and s == 1.0
    assert np.array_equal(t[0, :, :, 0], img[:, :, 2].astype(np.float32))              # BGR -> RGB
    # get_image: array passthrough is a copy, path loading, validation
    assert preprocess.get_image(img) is not img and np.array_equal(preprocess.get_image(img), img)
    p = str(tmp_path / "a.png")
    cv2.imwrite(p, img)
    assert np.array_equal(preprocess.get_image(p), img)
    with pytest.raises(ValueError, match="does not exist"):
        preprocess.get_image(str(tmp_path / "missing.png"))
    with pytest.raises(ValueError, match="Invalid image input"):
        preprocess.get_image(12345)
    import base64
    ok, buf = cv2.imencode(".png", img)
    uri = "data:image/png;base64," + base64.b64encode(buf.tobytes()).decode()
    assert np.array_equal(preprocess.load_base64_img(uri), img)                        # np.fromstring replaced (numpy 2)


def test_dropin_install_registers_reference_module_names():
    import sys
    from fire_b200 import dropin
    dropin.install("modules")
    try:
        import facenet_gpu
        import hnswlib
        from modules.encoder import Encoder
        from modules.hnsw_manager import HNSWManager
        assert facenet_gpu.FaceNetClient and hnswlib.Index and Encoder and HNSWManager
    finally:
        dropin.uninstall()
        sys.modules.pop("modules", None)
