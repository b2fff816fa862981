"""Three FaceNet512 forwards at B=256 (K1 preprocess + K2 + L2 norm): the target of the ncu captures in profiles/.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/ncu_forward.py
    ncu --set full --clock-control none --import-source on -k regex:block17 -c 1 -o prof python tools/ncu_forward.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch         # noqa: E402

from fire_b200 import _lib, engine, weights as W   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = engine.FaceNetEngine(512, W.synthetic_weights(512, 1234, calibrate=False))
crops = torch.randint(0, 256, (B, 160, 160, 3), dtype=torch.uint8, device="cuda")
boxes = torch.tensor([[0, 0, 160, 160]] * B, dtype=torch.int32, device="cuda")
fid = torch.arange(B, dtype=torch.int32, device="cuda")
desc = torch.tensor([[i * 76800, 160, 160, 480] for i in range(B)], dtype=torch.int64, device="cuda")
for _ in range(3):
    f16, _, _ = engine.preprocess_boxes(crops, desc, boxes, fid, _lib.PRE_REFERENCE, True, False)
    raw, l2 = eng.forward(f16, want_l2=True)
torch.cuda.synchronize()
print("ok", float(l2.abs().mean()))
