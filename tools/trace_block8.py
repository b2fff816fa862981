"""In-kernel timeline of the fused Block8 tails (block8_fused_kernel), FaceNet512 at batch B.

    FIRE_B200_TRACE8=1 python tools/trace_block8.py [B] 2> profiles/rNN_block8_timeline.txt
"""
import os
import sys

os.environ.setdefault("FIRE_B200_TRACE8", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch   # noqa: E402

from fire_b200 import engine, weights as W   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = engine.FaceNetEngine(512, W.synthetic_weights(512, 1234, calibrate=False))
x = engine.pixels_to_network_input(torch.randint(0, 256, (B, 160, 160, 3), device="cuda"))
for i in range(3):
    if i == 2:
        print("# ---- third forward (warm) ----", file=sys.stderr)
    eng.forward(x)
    torch.cuda.synchronize()
