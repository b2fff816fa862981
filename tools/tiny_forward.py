"""Smallest end-to-end invocation (for compute-sanitizer / debuggers): 3 n crops -> K1 -> FaceNet128 -> top-1 vs 300 rows.

    compute-sanitizer --tool memcheck python tools/tiny_forward.py [n]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

from fire_b200 import _lib, engine, weights as W   # noqa: E402

_lib.init(0)
frame = np.ascontiguousarray(np.tile(W.calibration_images(1, seed=3)[0], (2, 2, 1)))
flat, desc = engine.frames_to_device([frame])
boxes = torch.tensor([[10, 20, 160, 160], [-5, -5, 90, 70], [100, 50, 200, 240]] * (int(sys.argv[1]) if len(sys.argv) > 1 else 1), dtype=torch.int32).cuda()
f16, _, status = engine.preprocess_boxes(flat, desc, boxes, torch.zeros(len(boxes), dtype=torch.int32).cuda(), _lib.PRE_REFERENCE, True, False)
eng = engine.FaceNetEngine(128, W.synthetic_weights(128, 1234, calibrate=False))
raw, l2 = eng.forward(f16)
idx = engine.KnnIndex(128, 1000)
idx.add(torch.randn(300, 128, device="cuda"))
d, i = idx.search(l2.contiguous(), 1)
torch.cuda.synchronize()
print("tiny forward ok", raw.shape, float(raw.abs().mean()), i.flatten().tolist())
