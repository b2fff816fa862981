"""Writes tests/golden/facenet_golden.json from the CPU fp32 oracle (run here, committed; the GPU test reads it)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fire_b200 import weights as W            # noqa: E402
from oracle.facenet_ref import facenet_forward  # noqa: E402

SEED = 21
out = {"image_seed": SEED, "weights_seed": 1234, "note": "fp32 torch-CPU oracle on fire_b200.weights.calibration_images(4, seed)"}
for D in (128, 512):
    t = W.synthetic_weights(D, 1234)
    x = W.calibration_images(4, seed=SEED).astype(np.float32) / 255.0
    e = facenet_forward(t, x)
    out[str(D)] = {"embeddings": [[float(v) for v in row] for row in e]}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "facenet_golden.json"), "w") as f:
    json.dump(out, f)
print("written")
