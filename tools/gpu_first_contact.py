"""Staged bring-up of the sm_100a kernels on a real B200 (run under gpurun, each stage in its own
process with a timeout so a trap in one stage cannot take the others down).

    python tools/gpu_first_contact.py            # all stages
    python tools/gpu_first_contact.py knn_small  # one stage (used internally)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STAGES = ["knn_small", "knn_mid", "knn_k50", "pre", "conv_layers", "facenet_b2", "facenet_b32"]


def knn_case(N, D, Q, k, seed=0, margin=None):
    import numpy as np
    from fire_b200.engine import KnnIndex
    from oracle import native
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((N, D), dtype=np.float32)
    q = rng.standard_normal((Q, D), dtype=np.float32)
    idx = KnnIndex(D, capacity=N)
    if margin is not None:
        idx.set_margin(margin)
    idx.add(g)
    t = time.time(); dist, ids = idx.search(q, k); t = time.time() - t
    ora = native.BFIndexOracle(D); ora.add_items(g)
    ol, od = ora.knn_query(q, k, num_threads=8)
    same = (ids == ol.astype(np.int64))
    gap_ok = True
    bad = np.argwhere(~same)
    for (qi, j) in bad[:2000]:
        # a mismatch is tolerated only inside a 1e-5 tie window of the oracle
        lo, hi = max(0, j - 1), min(k - 1, j + 1)
        if not (abs(od[qi, j] - od[qi, lo]) < 1e-5 or abs(od[qi, j] - od[qi, hi]) < 1e-5):
            gap_ok = False
    print(f"knn N={N} D={D} Q={Q} k={k}: ids equal {same.mean():.6f}, mismatches {len(bad)}, tie-explained {gap_ok}, "
          f"max|dist diff| {abs(dist - od).max():.3e}, stats {idx.stats()}, first call {t*1e3:.1f} ms")
    assert gap_ok and abs(dist - od).max() < 5e-6
    return idx


def stage_knn_small():
    knn_case(1000, 128, 5, 1)
    knn_case(300, 128, 3, 10)
    knn_case(5000, 512, 130, 10)


def stage_knn_mid():
    knn_case(100000, 512, 300, 10)
    knn_case(200000, 128, 1000, 1)
    knn_case(20000, 512, 64, 10, margin=0.5)      # forces the exact fallback for every query


def stage_knn_k50():
    knn_case(30000, 512, 100, 50)
    knn_case(70, 128, 9, 50)


def stage_pre():
    import numpy as np, torch, cv2
    from fire_b200 import engine, _lib
    rng = np.random.default_rng(5)
    frame = cv2.GaussianBlur(rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8), (0, 0), 2)
    boxes = [[100, 100, 160, 160], [200, 50, 320, 320], [500, 300, 233, 201], [10, 10, 97, 83], [1800, 900, 400, 400],
             [-20, -30, 200, 180], [300, 300, 480, 320], [700, 200, 52, 47], [900, 500, 120, 300], [0, 0, 0, 10],
             [400, 400, 640, 480], [1000, 100, 161, 160], [64, 64, 480, 480], [5, 700, 333, 250]]
    flat, desc = engine.frames_to_device(frame[None])
    b = torch.tensor(boxes, dtype=torch.int32).cuda(); bf = torch.zeros(len(boxes), dtype=torch.int32).cuda()
    f16, f32, status = engine.preprocess_boxes(flat, desc, b, bf, _lib.PRE_REFERENCE, True, True)
    torch.cuda.synchronize()
    f16, f16_pad = engine.network_input_to_pixels(f16)
    f32 = f32.cpu().numpy(); f16 = f16.cpu().numpy(); f16_pad = f16_pad.cpu().numpy(); status = status.cpu().numpy()
    nbad = 0
    for i, (x, y, w, h) in enumerate(boxes):
        x, y, w, h = max(0, x), max(0, y), max(0, w), max(0, h)
        crop = frame[y:y + h, x:x + w]
        if crop.size == 0:
            assert status[i] == 1; continue
        ref = cv2.resize(crop, (160, 160), interpolation=cv2.INTER_AREA).astype(np.float32) / 255.0
        d = int((ref != f32[i]).sum()); nbad += d
        d16 = int((np.rint(ref * 255) != f16[i]).sum())
        print(f"pre box {boxes[i]} crop {crop.shape[:2]}: f32 mismatches {d}, f16 mismatches {d16}, pad nonzero {int((f16_pad[i] != 0).sum())}")
    assert nbad == 0


def stage_conv_layers():
    """Every op of the plan checked on its own against the CPU emulation of the same plan (no buffer reuse)."""
    import numpy as np, torch
    from fire_b200 import engine, weights as W
    from fire_b200.netplan import Plan
    import plan_emu
    D, B = 512, 3
    tensors = W.synthetic_weights(D, 1234)
    eng = engine.FaceNetEngine.__new__(engine.FaceNetEngine)
    # build an engine on a no-reuse plan so every intermediate survives the forward pass
    import ctypes as C
    from fire_b200 import _lib
    _lib.init(0)
    plan = Plan(D, fuse_siblings=True, reuse_buffers=False)
    blob = W.pack(plan, tensors)
    h = C.c_void_p(); buf = C.create_string_buffer(blob, len(blob))
    _lib.check(_lib.lib().fire_facenet_create(C.addressof(buf), len(blob), C.byref(h)))
    eng._h, eng.plan, eng.blob, eng.D, eng.device, eng._ws = h, plan, blob, D, torch.device("cuda", 0), None
    eng.num_ops = len(plan.ops)
    u8 = W.calibration_images(B, seed=11)
    xin = np.zeros((B, 160, 160, 8), np.float32); xin[..., :3] = u8
    ref_out, ref_bufs = plan_emu.run_plan(plan, blob, xin, honor_offsets=False, return_buffers=True)
    x = engine.pixels_to_network_input(torch.from_numpy(xin).cuda())
    raw, l2 = eng.forward(x)
    torch.cuda.synchronize()
    worst = 0.0
    for i, op in enumerate(plan.ops):
        if op.dst.buf == plan.out_buf:
            got = raw.cpu().numpy().reshape(B, 1, 1, D)
        else:
            got = eng.read_buffer(op.dst.buf, x)
        want = ref_bufs[op.dst.buf]
        sl = slice(op.dst.c_off, op.dst.c_off + op.dst.c)
        err = np.abs(got[..., sl] - want[..., sl]).max()
        scale = np.abs(want[..., sl]).max() + 1e-6
        flag = "" if err / scale < 2e-2 else "   <-- MISMATCH"
        worst = max(worst, err / scale)
        print(f"op {i:3d} {op.label:38s} k{op.kh}x{op.kw} s{op.stride} cin{op.cin:5d} cout{op.cout:5d} bn{op.bn_tile:4d} "
              f"max|err| {err:.4f} / scale {scale:.3f}{flag}")
    print("worst relative error over ops:", worst)
    assert worst < 2e-2


def _facenet(B):
    import numpy as np, torch
    from fire_b200 import engine, weights as W
    from oracle.facenet_ref import facenet_forward
    D = 512
    tensors = W.synthetic_weights(D, 1234)
    eng = engine.FaceNetEngine(D, tensors)
    rng = np.random.default_rng(2)
    u8 = np.concatenate([rng.integers(0, 256, (B // 2, 160, 160, 3), dtype=np.uint8), W.calibration_images(B - B // 2, seed=5)])
    x = torch.from_numpy(u8.astype(np.float32) / 255.0).cuda()
    raw, l2 = eng.encode_unit_f32(x)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(3):
        raw, l2 = eng.encode_unit_f32(x)
    torch.cuda.synchronize(); t = (time.time() - t) / 3
    ref = facenet_forward(tensors, u8.astype(np.float32) / 255.0)
    got = raw.cpu().numpy()
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    print(f"facenet512 B={B}: min cos {cos.min():.6f} mean {cos.mean():.6f}; l2 norm check {np.abs(np.linalg.norm(l2.cpu().numpy(), axis=1) - 1).max():.2e}; "
          f"{t*1e3:.2f} ms/forward = {B/t:.0f} embeds/s")
    assert cos.min() >= 0.9999


def stage_facenet_b2():
    _facenet(2)


def stage_facenet_b32():
    _facenet(32)
    _facenet(256)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        globals()["stage_" + sys.argv[1]]()
        print("STAGE OK", sys.argv[1])
        sys.exit(0)
    from fire_b200 import build
    from oracle import native
    build.build(); native.build()
    results = {}
    for s in STAGES:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], capture_output=True, text=True, timeout=420)
            out, rc = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() + (e.stderr or b"").decode() + "\nTIMEOUT", -9
        results[s] = rc
        print(f"===== stage {s}: rc={rc} ({time.time()-t:.1f}s)\n{out[-6000:]}", flush=True)
    print("SUMMARY", results)
    sys.exit(0 if all(v == 0 for v in results.values()) else 1)
