"""Per-op device time of one FaceNet forward (CUDA events around every launch, fire_facenet_profile).

    python tools/profile_ops.py [B] [D] > profiles/rNN_ops_B256.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

from fire_b200 import engine, weights as W   # noqa: E402
from fire_b200.netplan import OP_CONV        # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
eng = engine.FaceNetEngine(D, W.synthetic_weights(D, 1234, calibrate=False))
x = engine.pixels_to_network_input(torch.randint(0, 256, (B, 160, 160, 3), device="cuda"))
for _ in range(3):
    eng.forward(x)
ms = np.zeros(eng.num_ops)
reps = 5
for _ in range(reps):
    m, fl = eng.profile(x)
    ms += m
ms /= reps
print(f"# FaceNet{D} B={B}: {ms.sum():.3f} ms/forward (sum of per-op events) = {B / ms.sum() * 1e3:.0f} embeds/s; "
      f"{fl.sum() / ms.sum() / 1e9:.1f} TFLOP/s")
print(f"{'op':>3} {'label':38} {'M':>8} {'N':>5} {'K':>5} {'mode':>6} {'Mtiles':>7} {'ms':>8} {'TFLOP/s':>8} {'GB/s(min)':>9} {'%':>5}")
groups = {}
for i, op in enumerate(eng.plan.ops):
    M = B * op.Ho * op.Wo
    if op.kind == OP_CONV:
        mode = "tma" if (op.kh == 1 and op.kw == 1 and op.stride == 1) else "gather"
        grid = (M + 127) // 128
        tf = fl[i] / ms[i] / 1e9 if ms[i] > 0 else 0.0
        byts = 2 * (B * op.H * op.W * op.cin + M * op.cout + op.cout * op.k_pad)
        K = op.k_real
    else:
        mode, grid, tf, K = "pool", 0, 0.0, 0
        byts = 2 * (B * op.H * op.W * op.cin + M * op.cout)
    if op.kind == OP_CONV and fl[i] == 0:
        continue                      # part of a fused launch reported on its first op (block17_fused_kernel)
    print(f"{i:3d} {op.label:38} {M:8d} {op.cout:5d} {K:5d} {mode:>6} {grid:7d} {ms[i]:8.4f} {tf:8.1f} "
          f"{byts / ms[i] / 1e6:9.0f} {100 * ms[i] / ms.sum():5.1f}")
    key = op.label.split("_")[0] if not op.label.startswith("Conv2d") else "Stem"
    groups[key] = groups.get(key, 0.0) + ms[i]
print("# by stage:", {k: round(v, 3) for k, v in groups.items()})
