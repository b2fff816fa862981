"""Per-layer N-tile sweep: isolated per-op device time for candidate bn_tile values (FIRE_B200_BN override).

    python tools/tune_bn.py > gpurun_out/tune_bn.txt
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fire_b200.netplan import OP_CONV, Plan   # noqa: E402

plan = Plan(512)
# one representative op per layer class: (op index, candidate tiles)
classes = {}
for i, op in enumerate(plan.ops):
    if op.kind != OP_CONV:
        continue
    key = (op.cout, op.k_pad, op.Ho * op.Wo, bool(op.flags & 2), op.kh * op.kw > 1, op.stride)
    classes.setdefault(key, []).append(i)
cands = {}
for key, ops in classes.items():
    cout, residual = key[0], key[3]
    c = [d for d in range(16, 257, 16) if cout % d == 0 and (not residual or d % 64 == 0)]
    if len(c) > 1:
        cands[ops[0]] = (c, ops)
all_bn = sorted({d for c, _ in cands.values() for d in c})


def run(env_bn):
    env = dict(os.environ)
    if env_bn:
        env["FIRE_B200_BN"] = env_bn
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_ops.py"), "256", "512"], env=env, capture_output=True, text=True).stdout
    ms = {}
    for line in out.splitlines():
        f = line.split()
        if len(f) > 8 and f[0].isdigit():
            ms[int(f[0])] = float(f[7])
    return ms


base = run("")
results = {op: {"auto": base[op]} for op in cands}
for d in all_bn:
    spec = ",".join(f"{op}:{d}" for op, (c, _) in cands.items() if d in c)
    ms = run(spec)
    for op, (c, _) in cands.items():
        if d in c:
            results[op][d] = ms[op]
print("# isolated per-op ms (CUDA events, no PDL) by forced bn_tile; 'auto' = pick_bn_tile")
for op, (c, ops) in cands.items():
    o = plan.ops[op]
    best = min(((v, str(k)) for k, v in results[op].items()))
    row = "  ".join(f"{k}:{v:.4f}" for k, v in results[op].items())
    print(f"op {op:3d} {o.label:36s} x{len(ops):2d} M/img={o.Ho * o.Wo:5d} N={o.cout:4d} K={o.k_pad:4d}  best {best[1]} ({best[0]:.4f})  | {row}")
