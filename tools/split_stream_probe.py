"""Does running the two halves of a batch on two streams (two engine handles) hide the per-layer bubbles?

    python tools/split_stream_probe.py [B] [lanes...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch         # noqa: E402

from fire_b200 import engine, weights as W   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lanes_list = [int(a) for a in sys.argv[2:]] or [1, 2, 4]
tensors = W.synthetic_weights(512, 1234, calibrate=False)
x = engine.pixels_to_network_input(torch.randint(0, 256, (B, 160, 160, 3), device="cuda"))
for lanes in lanes_list:
    engs = [engine.FaceNetEngine(512, tensors) for _ in range(lanes)]
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    parts = list(x.chunk(lanes))
    raws = [torch.empty(p.shape[0], 512, device="cuda") for p in parts]
    l2s = [torch.empty(p.shape[0], 512, device="cuda") for p in parts]
    main = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event()
        ev.record(main)
        for e, s, p, r, l in zip(engs, streams, parts, raws, l2s):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                e.forward(p, want_l2=True, out_raw=r, out_l2=l)
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B} lanes={lanes}: {ms:.3f} ms/forward = {B / ms * 1e3:.0f} embeds/s")
    del engs
