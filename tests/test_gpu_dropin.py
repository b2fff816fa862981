"""GPU: the reference-facing surfaces (Encoder, HNSWManager, preprocess, dropin) end to end, against the all-CPU
chain cv2.resize(INTER_AREA)/255 -> fp32 oracle FaceNet -> L2 norm -> BFIndex oracle, including the reference's
accept/reject rule (face_recognition.py:412-469 restated in `_recognise`)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def encoder(fire_lib):
    os.environ["FIRE_B200_SYNTHETIC_WEIGHTS"] = "1"
    from fire_b200.encoder import Encoder
    return Encoder("512", "cpu_optimized")          # main.py:48 default mode string; ignored by design


def _frame(seed):
    from fire_b200 import weights as W
    tiles = W.calibration_images(12, seed=seed)     # 12 distinct structured 160x160 tiles -> one 480 x 640 frame
    rows = [np.concatenate(list(tiles[r * 4:(r + 1) * 4]), axis=1) for r in range(3)]
    return np.ascontiguousarray(np.concatenate(rows, axis=0))


def test_encoder_surface_matches_reference_arithmetic(encoder):
    import cv2
    from fire_b200 import weights as W
    from oracle.facenet_ref import facenet_forward
    assert encoder.input_shape == (160, 160) and encoder.output_shape == 512
    frame = _frame(3)
    crop = frame[37:291, 100:333]                                           # a non-contiguous BGR view, like image[y:y+h, x:x+w]
    pre = encoder.preprocess_for_encoder(crop)                              # modules/encoder.py:19-27
    want = np.expand_dims(cv2.resize(crop, (160, 160), interpolation=cv2.INTER_AREA).astype(np.float32) / 255.0, 0)
    assert pre.dtype == np.float32 and pre.shape == (1, 160, 160, 3) and np.array_equal(pre, want)
    emb = encoder.encode(pre)                                               # modules/encoder.py:16-17 -> facenet_gpu.py:127
    ref = facenet_forward(W.synthetic_weights(512, 1234), want)
    assert emb.shape == (1, 512) and emb.dtype == np.float32
    cos = float((emb * ref).sum() / (np.linalg.norm(emb) * np.linalg.norm(ref)))
    assert cos >= 0.9999
    for bad in (np.zeros((10, 10), np.uint8), np.zeros((10, 10, 4), np.uint8), np.zeros((0, 5, 3), np.uint8)):
        with pytest.raises(ValueError, match="incorrect shape"):
            encoder.preprocess_for_encoder(bad)
    # the batched additive path gives the same embeddings as the per-face reference path
    boxes = [[100, 37, 233, 254], [0, 0, 160, 160], [300, 200, 400, 400], [-10, -10, 100, 90]]
    batch, status = encoder.encode_crops([frame], boxes)
    assert list(status) == [0, 0, 0, 0]
    assert np.array_equal(batch[0], emb[0])


def _recognise(frame, boxes, preprocess, encode, query, labels, thr=0.7):
    """face_recognition.py:412-469 for faces without a track label, minus tracker / cache / enrolment."""
    out = []
    for (x, y, w, h) in boxes:
        x, y, w, h = max(0, x), max(0, y), max(0, w), max(0, h)
        face = frame[y:y + h, x:x + w]
        if face.size == 0:
            out.append(("skip", None)); continue
        e = encode(preprocess(face)).squeeze()
        n = np.linalg.norm(e)
        if n == 0:
            out.append(("skip", None)); continue
        e = e / n
        lab, dist = query(e, 1)
        cos = 1 - dist[0][0]
        out.append((labels[lab[0][0]], float(cos)) if cos > thr else ("Unknown", float(cos)))
    return out


def test_recognise_loop_decisions_match_cpu_chain(encoder, oracle_native, tmp_path):
    import cv2
    from fire_b200 import weights as W
    from fire_b200.hnsw_manager import HNSWManager
    from oracle.facenet_ref import FaceNetRef
    ref_net = FaceNetRef(W.synthetic_weights(512, 1234))
    frame, gallery_frame = _frame(11), _frame(12)
    boxes = [[0, 0, 160, 160], [160, 0, 160, 160], [320, 160, 160, 160], [100, 100, 200, 240], [480, 320, 400, 400],
             [-30, 200, 150, 180], [700, 10, 50, 50], [40, 300, 97, 83]]
    ref_pre = lambda f: np.expand_dims(cv2.resize(f, (160, 160), interpolation=cv2.INTER_AREA).astype(np.float32) / 255.0, 0)
    # gallery: 3 of the probe faces (exact enrolments), 3 faces of another frame, 4000 random identities
    rng = np.random.default_rng(5)
    enrol = [ref_net(ref_pre(frame[0:160, 0:160]))[0], ref_net(ref_pre(frame[0:160, 160:320]))[0],
             ref_net(ref_pre(frame[100:340, 100:300]))[0]] + \
            [ref_net(ref_pre(gallery_frame[0:160, i * 160:(i + 1) * 160]))[0] for i in range(3)]
    rows = np.concatenate([np.stack(enrol), rng.standard_normal((4000, 512)).astype(np.float32)])
    labels = [f"id{i}" for i in range(len(rows))]
    mgr = HNSWManager(512, str(tmp_path / "i"), str(tmp_path / "l"), str(tmp_path / "d"), None, max_elements=5000)
    for i, r in enumerate(rows[:6]):
        mgr.add_embedding(r, labels[i], i)                                   # one at a time, like the reference enrols
    mgr.load_embeddings_into_hnswlib([(i, labels[i], rows[i].tobytes()) for i in range(6, len(rows))])
    assert mgr.hnsw_index.get_current_count() == len(rows) and mgr.hnsw_labels == labels
    ora = oracle_native.BFIndexOracle(512)
    ora.add_items(rows)
    got = _recognise(frame, boxes, encoder.preprocess_for_encoder, encoder.encode, mgr.query, labels)
    want = _recognise(frame, boxes, ref_pre, ref_net, lambda e, k: ora.knn_query(e, k), labels)
    assert [g[0] for g in got] == [w[0] for w in want]                       # identical accept/reject + labels
    assert [w[0] for w in want][:4] == ["id0", "id1", "Unknown", "id2"] and want[6][0] == "skip"
    for g, w in zip(got, want):
        if w[1] is not None:
            assert abs(g[1] - w[1]) < 5e-3                                   # fp16 embeddings: cosines agree to the parity bar
    # persistence through the manager, then top-50 maintenance query
    mgr.save_hnswlib_index()
    mgr2 = HNSWManager(512, str(tmp_path / "i"), str(tmp_path / "l"), str(tmp_path / "d"), None, max_elements=5000)
    assert mgr2.hnsw_labels == labels and mgr2.hnsw_index.get_current_count() == len(rows)
    e0 = rows[0] / np.linalg.norm(rows[0])
    assert mgr2.find_similar_embeddings(e0, 0.999) == [0]
    l50, d50 = mgr2.query(e0, k=50)
    ol, od = ora.knn_query(e0, 50)
    assert np.array_equal(l50, ol) and np.abs(d50 - od).max() < 5e-6


def test_bulk_similar_lists_and_shrink_on_the_gpu_index(fire_lib, oracle_native, tmp_path):
    """SURVEY 8(f) row 3 on the real index: find_similar_all (tiled fire_knn_search_rows) == the per-row reference call, the
    bulk shrink_db_ids merges what those lists say, and the hnswlib-image save/load round trip keeps every answer."""
    import sqlite3
    from fire_b200.hnsw_manager import HNSWManager
    rng = np.random.default_rng(15)
    D, n = 512, 700
    centers = rng.standard_normal((60, D)).astype(np.float32)
    vecs = [(centers[rng.integers(0, 60)] + 0.25 * rng.standard_normal(D)).astype(np.float32) * float(rng.uniform(0.5, 2)) for _ in range(n)]
    labels = [f"person{i % 7}" if i % 5 == 0 else f"Unknown_{i:08x}" for i in range(n)]
    conn = sqlite3.connect(str(tmp_path / "f.db")); cur = conn.cursor()
    cur.execute("CREATE TABLE faces (id INTEGER PRIMARY KEY AUTOINCREMENT, label TEXT NOT NULL, embedding BLOB NOT NULL)")
    mgr = HNSWManager(D, str(tmp_path / "i"), str(tmp_path / "l"), str(tmp_path / "d"), None, max_elements=1000)
    for v, lab in zip(vecs, labels):
        cur.execute("INSERT INTO faces (label, embedding) VALUES (?, ?)", (lab, v.tobytes()))
        mgr.add_embedding(v, lab, cur.lastrowid)
    conn.commit()
    lists = mgr.find_similar_all(0.75, tile=256)
    ora = oracle_native.BFIndexOracle(D); ora.add_items(np.stack(vecs))
    for hid in range(0, n, 37):
        e = mgr._get_embedding_from_db_id(mgr.hnsw_db_ids[hid], cur)
        assert [int(x) for x in lists[hid]] == [int(x) for x in mgr.find_similar_embeddings(e, 0.75)]
        ol, od = ora.knn_query(e, 50)
        assert [int(x) for x in lists[hid]] == [int(l) for l, d in zip(ol[0], od[0]) if 1 - d >= 0.75]
    before = list(mgr.hnsw_labels)
    merges = mgr.shrink_db_ids(cur, conn, 0.75)
    assert merges > 0 and mgr.hnsw_labels != before
    assert [r[0] for r in cur.execute("SELECT label FROM faces ORDER BY id")] == mgr.hnsw_labels
    mgr2 = HNSWManager(D, str(tmp_path / "i"), str(tmp_path / "l"), str(tmp_path / "d"), None, max_elements=1000)   # the hnswlib image written by unify_labels
    assert mgr2.hnsw_labels == mgr.hnsw_labels and mgr2.hnsw_index.get_current_count() == n
    q = rng.standard_normal((5, D)).astype(np.float32)
    a, b = mgr.query_batch(q, k=10), mgr2.query_batch(q, k=10)
    assert np.array_equal(a[0], b[0]) and np.abs(a[1] - b[1]).max() < 2e-7


def test_preprocess_module_and_dropin(fire_lib):
    import sys
    import cv2
    from fire_b200 import dropin, preprocess
    frame = _frame(4)
    boxes = [[10, 20, 300, 200], [0, 0, 640, 480], [600, 400, 100, 100]]
    out, status = preprocess.crop_resize_normalize([frame], boxes, mode="reference")
    for i, (x, y, w, h) in enumerate(boxes):
        want = cv2.resize(frame[y:y + h, x:x + w], (160, 160), interpolation=cv2.INTER_AREA).astype(np.float32) / 255.0
        assert np.array_equal(out[i], want) and status[i] == 0
    ns, _ = preprocess.crop_resize_normalize([frame], boxes, mode="northstar")
    assert abs(float(ns[0].mean())) < 1e-4 and abs(float(ns[0].std()) - 1) < 1e-3
    with pytest.raises(ValueError):
        preprocess.crop_resize_normalize([frame], boxes, mode="bicubic")
    os.environ["FIRE_B200_SYNTHETIC_WEIGHTS"] = "1"
    dropin.install("engines")
    try:
        import facenet_gpu
        import hnswlib
        client = facenet_gpu.FaceNetClient(model_type=None, mode="gpu")      # main.py:45: default model type is None -> 128-d
        assert client.output_shape == 128 and client.model_name == "FaceNet-128d" and client.input_shape == (160, 160)
        e = client(np.zeros((2, 160, 160, 3), np.float32))
        assert e.shape == (2, 128) and np.array_equal(e[0], e[1])
        idx = hnswlib.Index(space="cosine", dim=128)
        idx.init_index(max_elements=100000, ef_construction=200, M=16)
        idx.set_ef(200)
        idx.add_items(e[0], 0)
        assert idx.get_current_count() == 1
        lab, dist = idx.knn_query(e[1], k=1)
        assert lab[0][0] == 0 and abs(dist[0][0]) < 1e-6
        with pytest.raises(RuntimeError):
            idx.knn_query(e[1], k=2)
    finally:
        dropin.uninstall()
        sys.modules.pop("modules", None)


def test_two_gpu_sharded_search_matches_single(fire_lib):
    """Runs only where >= 2 GPUs are visible (gpurun --gpus 2): torchrun, NCCL all_gather + fire_knn_merge."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "dist_knn_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_KNN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_batched_recognizer_on_the_gpu_engines(encoder, tmp_path):
    """fire_b200.recognizer with the real sm_100a surfaces (Encoder.encode_crops, HNSWManager.query_batch): the labels of
    a frame equal the per-face path (the orchestration logic itself is pinned against the reference's own loop on the
    CPU, tests/test_recognizer_host.py)."""
    import types
    from fire_b200 import recognizer
    from fire_b200.hnsw_manager import HNSWManager
    frame = _frame(6)
    boxes = [[0, 0, 160, 160], [160, 0, 160, 160], [320, 160, 160, 160], [480, 320, 160, 160], [700, 10, 50, 50]]
    mgr = HNSWManager(512, str(tmp_path / "i"), str(tmp_path / "l"), str(tmp_path / "d"), None, max_elements=1000)
    emb, _ = encoder.encode_crops([frame], boxes[:2])
    for j, lab in enumerate(("alice", "bob")):
        mgr.add_embedding(emb[j] / np.linalg.norm(emb[j]), lab, j)

    class FR:                                                   # the attributes recognize_faces touches (face_recognition.py:25-172)
        pass
    fr = FR()
    fr.start_time, fr.frame_index, fr.detection_interval, fr.frame_count = None, 0, 1, 0
    fr.total_detection_time = fr.total_encoding_time = 0.0
    fr.detect_faces = lambda image: [{'bbox': b, 'confidence': i} for i, b in enumerate(boxes)]
    fr.face_tracker = types.SimpleNamespace(update=lambda dets: [{'id': d['confidence'], 'bbox': d['bbox']} for d in dets])
    fr.track_id_to_label, fr.unknown_faces, fr.interested_label = {}, {}, None
    fr.encoder, fr.embedding_dim, fr.hnsw_manager, fr.similarity_threshold = encoder, 512, mgr, 0.7
    fr.recent_embeddings, fr.recent_labels = np.empty((0, 512), dtype=np.float32), []
    enrolled = []

    def handle_unknown(track_id, e, rename_label=None):
        lab = f"Unknown_{len(enrolled)}"
        enrolled.append(lab)
        mgr.add_embedding(e, lab, 100 + len(enrolled))
        return lab
    fr._handle_unknown_embedding = handle_unknown
    fr._add_to_recent_embeddings = lambda e, lab: (setattr(fr, "recent_embeddings", np.vstack([fr.recent_embeddings, e])), fr.recent_labels.append(lab))
    fr.update_label = lambda hid, new: None
    recognizer.install(fr)
    out = fr.recognize_faces(frame)
    assert [r['label'] for r in out] == ["alice", "bob", "Unknown_0", "Unknown_1"]          # the off-frame box is skipped
    assert out[0]['confidence'] > 0.9999 and mgr.hnsw_index.get_current_count() == 4
    again = fr.recognize_faces(frame)                            # tracked faces keep their labels without re-encoding
    assert [r['label'] for r in again] == ["alice", "bob", "Unknown_0", "Unknown_1"] and all(r['confidence'] == 1.0 for r in again)
