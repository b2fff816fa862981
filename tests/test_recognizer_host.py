"""SURVEY 8(f) row 4 - batched orchestration.  fire_b200.recognizer.recognize_faces_batched must leave a FaceRecognition
object in exactly the state the reference's own per-track loop (modules/face_recognition.py:371-489) leaves it in.
Runs on the CPU: the REAL reference class is imported from /root/reference (skipped where that tree is absent, e.g.
on the GPU box) with test doubles for the third-party pieces it needs (screeninfo, the SORT tracker, the detector,
the ONNX client, hnswlib -> fire_b200.hnswlib_compat over the BFIndex oracle)."""
import itertools
import os
import sys
import types

import numpy as np
import pytest

import fakes

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules")), reason="reference tree not present")
D = 128


class _FakeClient:
    """Deterministic stand-in for facenet_gpu.FaceNetClient: a fixed random projection of the 8x8-pooled crop."""
    input_shape, output_shape, model_name = (160, 160), D, "fake"

    def __init__(self, model_type="128", mode="gpu"):
        self.P = np.random.default_rng(42).standard_normal((8 * 8 * 3, D)).astype(np.float32)

    def __call__(self, img):
        x = np.asarray(img, dtype=np.float32).reshape(-1, 8, 20, 8, 20, 3).mean(axis=(2, 4)).reshape(-1, 192)
        return (x - 0.5) @ self.P


class _BatchEncoder:
    """What fire_b200.encoder.Encoder adds: encode_crops(frames, boxes) = the reference's per-face preprocess + encode."""

    def __init__(self, ref_encoder):
        self.ref = ref_encoder
        self.input_shape, self.output_shape = ref_encoder.input_shape, ref_encoder.output_shape
        self.preprocess_for_encoder, self.encode = ref_encoder.preprocess_for_encoder, ref_encoder.encode

    def encode_crops(self, frames, boxes, box_frame=None):
        out, status = np.zeros((len(boxes), D), np.float32), np.zeros(len(boxes), np.int32)
        for i, (x, y, w, h) in enumerate(boxes):
            x, y, w, h = max(0, x), max(0, y), max(0, w), max(0, h)
            crop = frames[0][y:y + h, x:x + w]
            if crop.size == 0:
                status[i] = 1
                continue
            out[i] = self.ref.encode(self.ref.preprocess_for_encoder(crop)).squeeze()
        return out, status


class _Tracker:
    """SORT stand-in: every detection keeps the id the script gives it."""

    def update(self, dets):
        return [{'id': d['confidence'], 'bbox': d['bbox']} for d in dets]        # the script smuggles the id in `confidence`


@pytest.fixture()
def ref_world(monkeypatch, tmp_path):
    from fire_b200 import hnswlib_compat
    monkeypatch.setattr(hnswlib_compat._engine, "KnnIndex", fakes.FakeKnnIndex)
    stubs = {
        "screeninfo": types.SimpleNamespace(get_monitors=lambda: []),
        "sort_UKF": types.SimpleNamespace(Sort=lambda **kw: _Tracker()),
        "yunet_face_detector": types.SimpleNamespace(detect_faces=lambda image: [], extract_faces=lambda *a, **k: []),
        "facenet_gpu": types.SimpleNamespace(FaceNetClient=_FakeClient),
        "hnswlib": hnswlib_compat,
    }
    for name, mod in stubs.items():
        m = types.ModuleType(name)
        m.__dict__.update(mod.__dict__)
        monkeypatch.setitem(sys.modules, name, m)
    for name in [n for n in sys.modules if n == "modules" or n.startswith("modules.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(REF)
    monkeypatch.chdir(tmp_path)
    import modules.face_recognition as frm
    yield frm
    for name in [n for n in sys.modules if n == "modules" or n.startswith("modules.")]:
        sys.modules.pop(name, None)


def _script(rng, n_frames):
    """Per frame: a list of (track id, bbox).  Ids persist for a few frames, boxes jitter, some are empty / off-frame."""
    people = [(int(rng.integers(0, 400)), int(rng.integers(0, 250)), int(rng.integers(50, 200)), int(rng.integers(50, 200))) for _ in range(9)]
    frames = []
    for f in range(n_frames):
        dets = []
        for pid, (x, y, w, h) in enumerate(people):
            if (f + pid) % 5 == 4:
                continue                                                   # the track drops out for a frame -> label forgotten
            tid = pid + 100 * ((f + pid) // 5)                             # and comes back under a new id
            dets.append((tid, [x + int(rng.integers(-3, 4)), y + int(rng.integers(-3, 4)), w, h]))
        if f % 4 == 1:
            dets.append((9000 + f, [700, 50, 40, 40]))                     # completely outside the 640 x 480 frame: empty crop
            dets.append((9500 + f, [-30, -20, 100, 90]))                   # negative corner: clamped, not shrunk
        frames.append(dets)
    return frames


def _run(frm, batched: bool, tmp, script, images, rename_at=None):
    import uuid as _uuid
    counter = itertools.count()
    frm.uuid.uuid4 = lambda: types.SimpleNamespace(hex=f"{next(counter):08x}" + "0" * 24)
    fr = frm.FaceRecognition(detector_type="yunet", encoder_model_type="128", encoder_mode="cpu_optimized", similarity_threshold=0.7,
                             unknown_trigger_count=1, sqlite_db_path=os.path.join(tmp, "f.db"), hnsw_index_path=os.path.join(tmp, "i.bin"),
                             hnsw_labels_path=os.path.join(tmp, "l.pkl"), hnsw_db_ids_path=os.path.join(tmp, "d.pkl"), max_recent=6,
                             detection_interval=1)
    if batched:
        from fire_b200 import recognizer
        fr.encoder = _BatchEncoder(fr.encoder)
        fr.hnsw_manager.query_batch = lambda E, k=1: fr.hnsw_manager.hnsw_index.knn_query(E, k=k)
        recognizer.install(fr)
    log = []
    for f, (dets, img) in enumerate(zip(script, images)):
        fr.detect_faces = lambda image, dets=dets: [{'bbox': b, 'confidence': tid} for tid, b in dets]
        log.append(fr.recognize_faces(img, rename_label="Renamed" if rename_at == f else None))
    state = dict(track=dict(fr.track_id_to_label), recent=fr.recent_embeddings.copy(), recent_labels=list(fr.recent_labels),
                 hnsw_labels=list(fr.hnsw_manager.hnsw_labels), count=fr.hnsw_manager.hnsw_index.get_current_count(),
                 unknown={k: (v['count'], np.array(v['embeddings'])) for k, v in fr.unknown_faces.items()}, frames=fr.frame_count)
    fr.close()
    return log, state


@pytest.mark.parametrize("rename_at", [None, 7])
def test_batched_recognizer_equals_the_reference_loop(ref_world, tmp_path, rename_at):
    rng = np.random.default_rng(5)
    script = _script(rng, 14)
    base = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    images = [np.ascontiguousarray(np.repeat(np.repeat(np.roll(base, f, axis=1), 8, axis=0), 8, axis=1)) for f in range(14)]
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    log_a, st_a = _run(ref_world, False, str(tmp_path / "a"), script, images, rename_at)
    log_b, st_b = _run(ref_world, True, str(tmp_path / "b"), script, images, rename_at)
    assert sum(len(r) for r in log_a) > 60 and st_a["count"] >= 3              # the scenario enrols, matches and re-identifies
    assert log_a == log_b
    assert st_a["track"] == st_b["track"] and st_a["recent_labels"] == st_b["recent_labels"] and st_a["hnsw_labels"] == st_b["hnsw_labels"]
    assert np.array_equal(st_a["recent"], st_b["recent"]) and st_a["count"] == st_b["count"] and st_a["frames"] == st_b["frames"]
    assert st_a["unknown"].keys() == st_b["unknown"].keys()
    for k in st_a["unknown"]:
        assert st_a["unknown"][k][0] == st_b["unknown"][k][0] and np.array_equal(st_a["unknown"][k][1], st_b["unknown"][k][1])


def test_bulk_shrink_db_ids_equals_the_reference_loop(ref_world, tmp_path):
    """HNSWManager.shrink_db_ids (all neighbour lists from tiled GPU passes, fire_knn_search_rows) leaves
    labels and SQLite exactly as the REAL FaceRecognition.shrink_db_ids (modules/face_recognition.py:265-315) does."""
    import sqlite3
    from fire_b200.hnsw_manager import HNSWManager
    frm = ref_world
    rng = np.random.default_rng(9)
    centers = rng.standard_normal((12, D)).astype(np.float32)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    people, vecs, labels = [], [], []
    for i in range(160):
        c = int(rng.integers(0, 12))
        v = centers[c] + float(rng.choice([0.02, 0.05, 0.12])) * rng.standard_normal(D).astype(np.float32)
        # clusters 0-7: one known name each plus unknowns; 8-9: two different known names (conflict); 10-11: unknowns only
        if c < 8:
            lab = f"person{c}" if rng.random() < 0.3 else f"Unknown_{i:08x}"
        elif c < 10:
            lab = f"person{c}{'ab'[int(rng.integers(0, 2))]}" if rng.random() < 0.5 else f"Unknown_{i:08x}"
        else:
            lab = f"Unknown_{i:08x}"
        vecs.append((v * float(rng.uniform(0.5, 3))).astype(np.float32)); labels.append(lab)   # un-normalised on purpose (M2)

    (tmp_path / "ref").mkdir(); (tmp_path / "new").mkdir()
    fr = frm.FaceRecognition(detector_type="yunet", encoder_model_type="128", encoder_mode="cpu_optimized",
                             sqlite_db_path=str(tmp_path / "ref" / "f.db"), hnsw_index_path=str(tmp_path / "ref" / "i.bin"),
                             hnsw_labels_path=str(tmp_path / "ref" / "l.pkl"), hnsw_db_ids_path=str(tmp_path / "ref" / "d.pkl"))
    for v, lab in zip(vecs, labels):
        fr.hnsw_manager.add_embedding(v, lab, fr._add_to_sqlite(lab, v))
    fr.shrink_db_ids(0.75)
    want_labels = list(fr.hnsw_manager.hnsw_labels)
    want_db = [r[0] for r in fr.db_manager.cursor.execute("SELECT label FROM faces ORDER BY id")]
    fr.close()

    conn = sqlite3.connect(str(tmp_path / "new" / "f.db"))
    cur = conn.cursor()
    cur.execute("CREATE TABLE faces (id INTEGER PRIMARY KEY AUTOINCREMENT, label TEXT NOT NULL, embedding BLOB NOT NULL)")
    mgr = HNSWManager(D, str(tmp_path / "new" / "i.bin"), str(tmp_path / "new" / "l.pkl"), str(tmp_path / "new" / "d.pkl"), None)
    for v, lab in zip(vecs, labels):
        cur.execute("INSERT INTO faces (label, embedding) VALUES (?, ?)", (lab, v.tobytes()))
        mgr.add_embedding(v, lab, cur.lastrowid)
    conn.commit()
    lists = mgr.find_similar_all(0.75, tile=64)
    for hid in (0, 17, 159):                                                   # the bulk lists == the per-row call of the reference API
        e = mgr._get_embedding_from_db_id(mgr.hnsw_db_ids[hid], cur)
        assert [int(x) for x in lists[hid]] == [int(x) for x in mgr.find_similar_embeddings(e, 0.75)]
    merges = mgr.shrink_db_ids(cur, conn, 0.75)
    assert merges >= 8 and len(set(want_labels)) < len(set(labels))           # the scenario really merges groups
    assert mgr.hnsw_labels == want_labels
    assert [r[0] for r in cur.execute("SELECT label FROM faces ORDER BY id")] == want_db
    assert any(l.startswith("person8") for l in want_labels) and len({l for l in want_labels if l.startswith("person8")}) == 2   # conflict kept
