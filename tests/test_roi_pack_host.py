"""Host side of the ROI upload (fire_pack_rois_host, configs[4]): the packed rectangles + tables describe exactly the crops
the reference slices on the host (modules/face_recognition.py:412-420), for boxes that cross the frame edges, start at
negative coordinates or miss the frame entirely.  No GPU involved: this entry point only stages bytes for the upload."""
import numpy as np


def test_packed_rois_are_the_reference_crops(fire_lib):
    from fire_b200 import engine
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((108, 192), (77, 130), (300, 40))]
    flat = np.concatenate([f.reshape(-1) for f in frames])
    desc, off = [], 0
    for f in frames:
        desc.append((off, f.shape[0], f.shape[1], f.shape[1] * 3)); off += f.size
    desc = np.array(desc, dtype=np.int64)
    boxes = np.array([[10, 20, 50, 40], [-5, -7, 30, 30], [180, 100, 40, 40], [0, 0, 192, 108], [500, 10, 20, 20], [3, 3, 0, 9],
                      [100, 60, 64, 33], [-50, 10, 20, 20], [7, 250, 33, 100], [0, 0, 1, 1]], dtype=np.int32)
    bframe = np.array([0, 0, 0, 0, 1, 1, 1, 2, 2, 2], dtype=np.int32)
    out = np.zeros(1 << 20, dtype=np.uint8)
    out = out[(-out.ctypes.data) % 16:]
    for threads in (1, 4):
        used = engine.pack_rois(flat, desc, boxes, bframe, out, threads=threads)
        n = len(boxes)
        assert used % 256 == 0 and used <= out.nbytes
        d = out[:32 * n].view(np.int64).reshape(n, 4)
        b = out[32 * n:48 * n].view(np.int32).reshape(n, 4)
        bf = out[48 * n:52 * n].view(np.int32)
        assert np.array_equal(bf, np.arange(n))
        for i, (x, y, w, h) in enumerate(boxes):
            x, y, w, h = max(0, x), max(0, y), max(0, w), max(0, h)            # the reference's independent clamps
            want = frames[bframe[i]][y:y + h, x:x + w]                           # numpy clips the far edges
            o, rows, cols, pitch = (int(v) for v in d[i])
            if want.size == 0:
                assert rows == 0 and cols == 0 and list(b[i]) == [0, 0, 0, 0]
                continue
            assert (rows, cols) == want.shape[:2] and pitch % 16 == 0 and o % 256 == 0 and list(b[i]) == [0, 0, cols, rows]
            got = np.stack([out[o + r * pitch:o + r * pitch + cols * 3] for r in range(rows)]).reshape(rows, cols, 3)
            assert np.array_equal(got, want)
