"""Small kNN workload for profilers: 1M x 512 gallery, 4096 queries, top-10, a few searches.

    python tools/knn_probe.py [N] [Q] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch   # noqa: E402

from fire_b200.engine import KnnIndex   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda"); g.manual_seed(3)
idx = KnnIndex(512, N)
idx.add(torch.randn(N, 512, generator=g, device="cuda"))
q = torch.randn(Q, 512, generator=g, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
idx.search(q, 10)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    d, i = idx.search(q, 10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"N={N} Q={Q}: {ms:.3f} ms/batch = {Q / ms * 1e3:.0f} QPS, {2.0 * Q * N * 512 / ms / 1e9:.1f} TFLOP/s, stats {idx.stats()}")
