"""fire_b200.onnx_reader: parses ONNX wire format by hand and maps it onto the plan's Keras tensor names."""
import numpy as np
import pytest

import onnx_writer
from fire_b200 import onnx_reader, weights as W
from fire_b200.netplan import Plan


@pytest.mark.parametrize("fold_bn", [False, True])
def test_roundtrip_through_an_onnx_file(tmp_path, fold_bn):
    t = W.synthetic_weights(128, 7, calibrate=False)
    path = str(tmp_path / "facenet128.onnx")
    onnx_writer.write_facenet_like(path, t, fold_bn=fold_bn)
    nodes, inits = onnx_reader.parse_model(path)
    assert sum(n["op"] == "Conv" for n in nodes) == 132 and sum(n["op"] == "MatMul" for n in nodes) == 1
    got = onnx_reader.load_facenet_tensors(path, 128)
    assert set(got) == set(t)
    plan = Plan(128)
    blob_a, blob_b = W.pack(plan, t), W.pack(Plan(128), got)
    if not fold_bn:
        for k in t:
            assert np.array_equal(got[k], t[k]), k
        assert blob_a == blob_b
    else:                                  # BN folded into the conv by the exporter: the folded weights/bias must agree
        a = np.frombuffer(blob_a, dtype=np.uint16)
        b = np.frombuffer(blob_b, dtype=np.uint16)
        assert a.shape == b.shape and (a != b).mean() < 0.02        # identical up to 1-ulp fp16 re-rounding


def test_lfs_pointer_is_rejected(tmp_path):
    p = tmp_path / "facenet512.onnx"
    p.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:f0dfb218\nsize 94037431\n")
    with pytest.raises(ValueError, match="git-LFS pointer"):
        onnx_reader.parse_model(str(p))
