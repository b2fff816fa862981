"""FaceNet512 forward latency at small batches (the reference encodes ONE face per call, modules/encoder.py:26), with the
fused chains on / off: which launch structure is the faster one when a batch cannot fill the GPU.

    python tools/small_batch_probe.py            # prints ms per forward for B = 1, 4, 8, 16, 32, 64
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch   # noqa: E402

from fire_b200 import engine, weights as W   # noqa: E402

t = W.synthetic_weights(512, 1234, calibrate=False)
cfgs = [("default", {}), ("FUSE17=0", {"FIRE_B200_FUSE17": "0"}), ("FUSE35=0", {"FIRE_B200_FUSE35": "0"}), ("FUSE8=0", {"FIRE_B200_FUSE8": "0"}),
        ("all per layer", {"FIRE_B200_FUSE17": "0", "FIRE_B200_FUSE35": "0", "FIRE_B200_FUSE8": "0"}), ("PDL=0", {"FIRE_B200_PDL": "0"})]
for name, env in cfgs:
    for k in ("FIRE_B200_FUSE17", "FIRE_B200_FUSE35", "FIRE_B200_FUSE8", "FIRE_B200_PDL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = engine.FaceNetEngine(512, t)
    row = []
    for B in (1, 4, 8, 16, 32, 64):
        x = engine.pixels_to_network_input(torch.randint(0, 256, (B, 160, 160, 3), device="cuda"))
        for _ in range(5):
            eng.forward(x)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            eng.forward(x)
        b.record()
        torch.cuda.synchronize()
        row.append(a.elapsed_time(b) / 50)
    print(f"{name:14s} launches {eng.num_launches:3d} | " + "  ".join(f"B={B}: {v:.3f} ms" for B, v in zip((1, 4, 8, 16, 32, 64), row)), flush=True)
    eng.close()
