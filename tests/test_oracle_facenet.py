"""Structural pins for the FaceNet oracle (no reference output exists to pin values against: PARITY UNPINNED)."""
import numpy as np
import torch

from oracle.facenet_ref import expected_param_count, facenet_forward, weight_shapes


def test_parameter_count_matches_lfs_pointer_sizes():
    """weights/facenet128.onnx / facenet512.onnx LFS pointers record 91 256 008 / 94 037 431 bytes;
    the fp32 parameters of this graph account for them to 0.06 %, which pins 3-parameter BN and bias-only-on-up."""
    for D, file_bytes in ((128, 91_256_008), (512, 94_037_431)):
        n = expected_param_count(D)
        assert n == {128: 22_808_144, 512: 23_497_424}[D]
        assert 0 < file_bytes - 4 * n < 0.0006 * file_bytes


def test_conv_count_and_shapes():
    s = weight_shapes(512)
    kernels = [k for k in s if k.endswith("/kernel")]
    assert len(kernels) == 133                                        # 132 convs + the bottleneck Dense
    assert sum(1 for k in s if k.endswith("/bias")) == 21             # the 21 scaled residual "up" convs
    assert s["Bottleneck/kernel"] == (1792, 512) and s["Conv2d_1a_3x3/kernel"] == (3, 3, 3, 32)


def test_fp32_agrees_with_fp64_and_is_not_l2_normalised():
    from fire_b200 import weights as W
    t = W.synthetic_weights(128, 1234)
    x = W.calibration_images(3, seed=8).astype(np.float32) / 255.0
    a = facenet_forward(t, x)
    b = facenet_forward(t, x, dtype=torch.float64)
    cos = (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    assert a.shape == (3, 128) and cos.min() > 0.999999
    assert np.all(np.abs(np.linalg.norm(a, axis=1) - 1) > 0.1)        # encode() returns the raw embedding
    one = facenet_forward(t, x[1:2])
    assert np.abs(one[0] - a[1]).max() < 1e-3 * np.abs(a[1]).max()    # batch independent (up to oneDNN blocking)
