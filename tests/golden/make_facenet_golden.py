"""Writes tests/golden/facenet_golden.json from an executor NEITHER the oracle NOR the product wrote: the synthetic weights
are exported as a full-structure ONNX file (tests/onnx_graph_writer.py: branches, Concat, the `scaling` Mul + Add, BN
nodes) and run with cv2.dnn.readNetFromONNX - the one ONNX runtime available in this image (SURVEY 8c lists it as the
independent second opinion).  The CPU oracle (tests/test_oracle_facenet.py) and the GPU engine
(tests/test_gpu_facenet.py::test_golden_embeddings) are both checked against these numbers.

    python tests/golden/make_facenet_golden.py      (run in the build container; the JSON is committed)
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import cv2                                     # noqa: E402
import onnx_graph_writer                       # noqa: E402
from fire_b200 import weights as W            # noqa: E402

SEED = 21


def dnn_embeddings(tensors, D, x_nhwc, fold_bn=False):
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, f"facenet{D}.onnx")
        onnx_graph_writer.write_facenet_graph(path, tensors, D, fold_bn=fold_bn)
        net = cv2.dnn.readNetFromONNX(path)
        net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
        net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)
        outs = []
        for i in range(x_nhwc.shape[0]):           # batch 1 per call, like the reference (modules/encoder.py:26)
            net.setInput(np.ascontiguousarray(x_nhwc[i:i + 1]))
            outs.append(net.forward().copy())
    return np.concatenate(outs).astype(np.float32)


if __name__ == "__main__":
    out = {"image_seed": SEED, "weights_seed": 1234, "executor": f"cv2.dnn.readNetFromONNX, OpenCV {cv2.__version__}, CPU target",
           "note": "full-structure ONNX export of fire_b200.weights.synthetic_weights(D, 1234) run on fire_b200.weights.calibration_images(4, seed)"}
    for D in (128, 512):
        t = W.synthetic_weights(D, 1234)
        x = W.calibration_images(4, seed=SEED).astype(np.float32) / 255.0
        e = dnn_embeddings(t, D, x)
        out[str(D)] = {"embeddings": [[float(v) for v in row] for row in e]}
    with open(os.path.join(HERE, "facenet_golden.json"), "w") as f:
        json.dump(out, f)
    print("written")
