"""Batched orchestration of the identification hot path.

The reference encodes ONE face per `session.run` and issues ONE `knn_query` per face inside its per-track loop
(modules/face_recognition.py:404-486).  `recognize_faces_batched` produces exactly the same results and side effects
for a `FaceRecognition`-shaped object, but runs the expensive part once per frame:

    1. collect every track that has no label yet and a non-empty crop      (face_recognition.py:408-420)
    2. ONE crop/resize/normalise + FaceNet launch chain for all of them    (Encoder.encode_crops -> K1 + K2 on the B200)
    3. ONE batched cosine top-1 against the gallery as it is before the frame (HNSWManager.query_batch -> K3)
    4. the decision loop, in the reference's track order: recent-embedding cache, threshold, unknown handling,
       cache append (face_recognition.py:446-477).  Steps 4's side effects can change the gallery (an unknown face is
       enrolled, a label is renamed); from that point on the remaining faces of the frame are re-queried live, so the
       outcome is identical to the sequential reference, just with fewer device round trips.

`install(fr)` rebinds `fr.recognize_faces`; everything else of the caller (detector, tracker, SQLite, annotate) is
untouched.  Requires the fire_b200 drop-in surfaces: `fr.encoder.encode_crops` (fire_b200.encoder.Encoder) and
`fr.hnsw_manager.query_batch` (fire_b200.hnsw_manager.HNSWManager).
"""
from __future__ import annotations

import logging
import time
import types

import numpy as np


def _clamped_crop_shape(image: np.ndarray, bbox):
    """The reference's crop rule: x, y, w, h are clamped to >= 0 independently, numpy clips the far edge."""
    x, y, w, h = (max(0, int(v)) for v in bbox)
    H, W = image.shape[:2]
    return max(0, min(H, y + h) - min(H, y)), max(0, min(W, x + w) - min(W, x))


def recognize_faces_batched(fr, image: np.ndarray, rename_label: str = None):
    """Same contract as FaceRecognition.recognize_faces(image, rename_label) (face_recognition.py:371-489)."""
    if fr.start_time is None:
        fr.start_time = time.time()
    fr.frame_index += 1

    # detection / tracking cadence (face_recognition.py:378-393)
    if fr.frame_index % fr.detection_interval == 0:
        t0 = time.time()
        found = fr.detect_faces(image)
        fr.total_detection_time += time.time() - t0
        tracks = fr.face_tracker.update([{'bbox': d.get('bbox', [0, 0, 0, 0]), 'confidence': d.get('confidence', 1.0)} for d in found])
    else:
        tracks = fr.face_tracker.update([])

    # forget tracks that disappeared (face_recognition.py:395-401)
    alive = {t['id'] for t in tracks}
    for tid in set(fr.track_id_to_label) - alive:
        del fr.track_id_to_label[tid]
        fr.unknown_faces.pop(tid, None)

    # ---- 1. which tracks need an embedding ------------------------------------------------------------------
    pending = []                       # indices into `tracks`
    for i, trk in enumerate(tracks):
        if trk['id'] in fr.track_id_to_label:
            continue
        ch, cw = _clamped_crop_shape(image, trk['bbox'])
        if ch == 0 or cw == 0:
            logging.warning(f"Face image has zero size for track ID {trk['id']}. Skipping.")
            continue
        pending.append(i)

    # ---- 2. + 3. one encode, one gallery query -------------------------------------------------------------------
    unit = {}                          # track index -> unit-norm embedding (None: the reference would skip this face)
    pre_query = {}                     # track index -> (labels[1,1], distances[1,1]) against the gallery before this frame
    if pending:
        t0 = time.time()
        raw, status = fr.encoder.encode_crops([image], [tracks[i]['bbox'] for i in pending])
        fr.total_encoding_time += time.time() - t0
        rows = []
        for j, i in enumerate(pending):
            e = np.asarray(raw[j])
            if status[j] != 0:
                unit[i] = None
                continue
            if e.shape[0] != fr.embedding_dim:
                logging.error(f"Invalid embedding size: expected {fr.embedding_dim}, got {e.shape[0]}")
                unit[i] = None
                continue
            n = np.linalg.norm(e)
            if n == 0:
                logging.error("Received zero vector from encoder. Skipping this face.")
                unit[i] = None
                continue
            unit[i] = e / n
            rows.append(i)
        if rows and fr.hnsw_manager.hnsw_index.get_current_count() > 0:
            labels, dists = fr.hnsw_manager.query_batch(np.stack([unit[i] for i in rows]), k=1)
            if labels is not None:
                for j, i in enumerate(rows):
                    pre_query[i] = (labels[j:j + 1], dists[j:j + 1])
    gallery_version = (fr.hnsw_manager.hnsw_id_counter, id(fr.hnsw_manager.hnsw_labels), tuple(fr.hnsw_manager.hnsw_labels[-1:]))

    # ---- 4. decisions, in the reference's order --------------------------------------------------------------------
    results = []
    gallery_dirty = False
    for i, trk in enumerate(tracks):
        tid, bbox = trk['id'], trk['bbox']
        if tid in fr.track_id_to_label:
            label, confidence = fr.track_id_to_label[tid], 1.0
        else:
            e = unit.get(i)
            if e is None:
                continue
            label, confidence = "Unknown", 0.0
            if fr.recent_embeddings.shape[0] > 0:                                   # recent cache first (:450-456)
                sims = np.dot(fr.recent_embeddings, e.T).flatten()
                best = int(np.argmax(sims))
                if sims[best] > fr.similarity_threshold:
                    label, confidence = fr.recent_labels[best], float(sims[best])
            if label == "Unknown":                                                  # gallery (:459-469)
                if gallery_dirty:
                    labels, dists = fr.hnsw_manager.query(e, k=1)
                else:
                    labels, dists = pre_query.get(i, (None, None))
                if labels is not None and labels.size > 0:
                    cos = 1 - dists[0][0]
                    if cos > fr.similarity_threshold:
                        hid = labels[0][0]
                        label, confidence = fr.hnsw_manager.hnsw_labels[hid], float(cos)
                        if rename_label:
                            fr.update_label(hid, rename_label)
                            label = rename_label
                            gallery_dirty = True
            if label == "Unknown":                                                  # unknown handling may enrol (:472-474)
                label = fr._handle_unknown_embedding(tid, e, rename_label)
                confidence = 1.0
                if (fr.hnsw_manager.hnsw_id_counter, id(fr.hnsw_manager.hnsw_labels), tuple(fr.hnsw_manager.hnsw_labels[-1:])) != gallery_version:
                    gallery_dirty = True
            fr.track_id_to_label[tid] = label
            fr._add_to_recent_embeddings(e, label)
        if fr.interested_label is not None and label != fr.interested_label:
            continue
        results.append({'label': fr.track_id_to_label[tid], 'confidence': float(confidence), 'bbox': bbox})

    fr.frame_count += 1
    return results


def install(fr):
    """Rebind fr.recognize_faces to the batched implementation (process_video / process_webcam call it per frame)."""
    for need, where in (("encode_crops", fr.encoder), ("query_batch", fr.hnsw_manager)):
        if not hasattr(where, need):
            raise TypeError(f"{type(where).__name__} has no {need}(): install the fire_b200 drop-in at level 'modules' first")
    fr.recognize_faces = types.MethodType(recognize_faces_batched, fr)
    return fr
