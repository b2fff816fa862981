"""fire_b200.onnx_reader: parses ONNX wire format by hand and maps it onto the plan's Keras tensor names."""
import numpy as np
import pytest

import onnx_graph_writer
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
    with pytest.raises(ValueError, match="Concat"):            # a linear chain is not the FaceNet graph: the structure check says so
        onnx_reader.load_facenet_tensors(path, 128)
    got = onnx_reader.load_facenet_tensors(path, 128, verify=False)
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


@pytest.mark.parametrize("D,fold_bn", [(128, False), (512, True)])
def test_full_structure_graph(tmp_path, D, fold_bn):
    """E2: a file with the REAL structure (three branch families, Concat, the `scaling` Mul + Add, BN nodes or BN folded by the
    exporter, the Dense BN decomposed into Mul + Add) - the same file cv2.dnn executes in tests/test_oracle_facenet.py - maps
    onto the plan's tensors, and the residual scales are read from the graph and checked."""
    from oracle.facenet_ref import facenet_forward
    t = W.synthetic_weights(D, 7, calibrate=False)
    path = str(tmp_path / f"facenet{D}.onnx")
    onnx_graph_writer.write_facenet_graph(path, t, D, fold_bn=fold_bn)
    nodes, _ = onnx_reader.parse_model(path)
    ops = [n["op"] for n in nodes]
    assert ops.count("Conv") == 132 and ops.count("Concat") == 23 and ops.count("Mul") == 22 and ops.count("Add") == 22
    got = onnx_reader.load_facenet_tensors(path, D)
    assert set(got) == set(t)
    if not fold_bn:                                            # every conv tensor verbatim; only the decomposed Dense BN is re-expressed
        assert [k for k in t if not np.array_equal(got[k], t[k])] == [f"Bottleneck_BatchNorm/{q}" for q in ("beta", "moving_mean", "moving_variance")]
    x = W.calibration_images(2, seed=3).astype(np.float32) / 255.0
    a, b = facenet_forward(t, x), facenet_forward(got, x)
    assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()       # the same network


def test_wrong_residual_scale_is_rejected(tmp_path, monkeypatch):
    t = W.synthetic_weights(128, 7, calibrate=False)
    path = str(tmp_path / "facenet128.onnx")
    real = onnx_graph_writer._G.residual

    def tampered(self, x, branches, p, scale, relu):
        return real(self, x, branches, p, 0.3 if p == "Block17_4" else scale, relu)
    monkeypatch.setattr(onnx_graph_writer._G, "residual", tampered)
    onnx_graph_writer.write_facenet_graph(path, t, 128)
    with pytest.raises(ValueError, match="Block17_4.*residual scale"):
        onnx_reader.load_facenet_tensors(path, 128)
