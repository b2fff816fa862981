"""In-kernel timeline of every conv layer of ONE real forward (PDL on), via FIRE_B200_TRACE_ALL=1.

    FIRE_B200_TRACE_ALL=1 python tools/trace_forward.py [B] 2> profiles/rNN_forward_timeline.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FIRE_B200_TRACE_ALL"] = "1"
import torch         # noqa: E402

from fire_b200 import engine, weights as W   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = engine.FaceNetEngine(512, W.synthetic_weights(512, 1234, calibrate=False))
x = engine.pixels_to_network_input(torch.randint(0, 256, (B, 160, 160, 3), device="cuda"))
for i in range(3):
    if i == 2:
        print("# ---- third forward (warm) ----", file=sys.stderr)
    eng.forward(x)
