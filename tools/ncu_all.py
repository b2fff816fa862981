"""One pass over every kernel family of the hot path, for `ncu --set full` (profiles/r02_ncu_*): no warm-up, every workload once.

    K1   preprocess_reference_kernel: 256 x 160x160 crops (copy case) and 256 configs[4]-sized boxes out of 1080p frames
    K2   one FaceNet512 forward at B=256 (30 tensor launches + 2 max-pools + L2 norm)
    K3   exact top-10 over 1M x 512: Q = 4096 (tensor-bound) and Q = 1, 32, 256 (HBM-bound small batches)

    python tools/ncu_all.py && ncu --set full --clock-control none -o gpurun_out/r02_full python tools/ncu_all.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

from fire_b200 import _lib, engine, weights as W   # noqa: E402
from fire_b200.engine import KnnIndex               # noqa: E402

B = 256
dev = "cuda"
eng = engine.FaceNetEngine(512, W.synthetic_weights(512, 1234, calibrate=False))
crops = torch.randint(0, 256, (B, 160, 160, 3), dtype=torch.uint8, device=dev)
boxes = torch.tensor([[0, 0, 160, 160]] * B, dtype=torch.int32, device=dev)
fid = torch.arange(B, dtype=torch.int32, device=dev)
desc = torch.tensor([[i * 76800, 160, 160, 480] for i in range(B)], dtype=torch.int64, device=dev)
rng = np.random.default_rng(5)
fr4 = torch.randint(0, 256, (8, 1080, 1920, 3), dtype=torch.uint8, device=dev)
bx4 = np.stack([rng.integers(0, 1500, B), rng.integers(0, 660, B), rng.integers(48, 401, B), rng.integers(48, 401, B)], 1).astype(np.int32)
d4 = torch.tensor([[i * 1080 * 1920 * 3, 1080, 1920, 5760] for i in range(8)], dtype=torch.int64, device=dev)
bx4_t, bf4 = torch.from_numpy(bx4).to(dev), torch.arange(B, dtype=torch.int32, device=dev) % 8
g = torch.Generator(device=dev); g.manual_seed(3)
idx = KnnIndex(512, 1_000_000)
gal = torch.randn(1_000_000, 512, generator=g, device=dev)
q = torch.randn(4096, 512, generator=g, device=dev)
torch.cuda.synchronize()

# ---- profiled region ------------------------------------------------------------------------------------------------
engine.preprocess_boxes(fr4, d4, bx4_t, bf4, _lib.PRE_REFERENCE, True, False)              # K1, configs[4] boxes
f16, _, _ = engine.preprocess_boxes(crops, desc, boxes, fid, _lib.PRE_REFERENCE, True, False)   # K1, copy case
raw, l2 = eng.forward(f16, want_l2=True)                                                    # K2
idx.add(gal)                                                                                # knn_normalize_kernel (enrol)
for Q in (4096, 1, 32, 256):
    idx.search(q[:Q].contiguous(), 10)                                                      # K3
torch.cuda.synchronize()
print("ok", float(l2.abs().mean()), "algorithmic bytes K1 boxes", int((bx4[:, 2].astype(np.int64) * bx4[:, 3] * 3).sum() + B * 80 * 80 * 16 * 2))
