"""Phase timeline of block35_fused_kernel inside ONE real forward (FIRE_B200_TRACE35=1), plus its event time.

    python tools/trace_block35.py [B] 2> profiles/rNN_block35_timeline.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FIRE_B200_TRACE35"] = "1"
import torch         # noqa: E402

from fire_b200 import engine, weights as W   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = engine.FaceNetEngine(512, W.synthetic_weights(512, 1234, calibrate=False))
x = engine.pixels_to_network_input(torch.randint(0, 256, (B, 160, 160, 3), device="cuda"))
for i in range(3):
    if i == 2:
        print("# ---- third forward (warm) ----", file=sys.stderr)
    eng.forward(x)
ms, fl = eng.profile(x)
first = next(i for i, op in enumerate(eng.plan.ops) if op.label == "Block35_1_heads")
print(f"# block35_fused_kernel alone (events): {ms[first]:.4f} ms, {fl[first] / ms[first] / 1e9:.1f} TFLOP/s", file=sys.stderr)
