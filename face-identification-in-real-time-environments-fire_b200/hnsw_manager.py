"""Drop-in for the reference's ``modules/hnsw_manager.py`` (modules/hnsw_manager.py:11-262) over the B200 exact index.

The class keeps the reference's constructor, its public attributes (`hnsw_index`, `hnsw_labels`, `hnsw_db_ids`,
`hnsw_id_counter`, the three paths and `encryptor`) and every method name and signature
(`tests/golden/reference_api.json` pins them), because modules/face_recognition.py reaches into all of them
(SURVEY §8a rows M1-M6).  The observable rules are the reference's, including the ones a rewrite would be tempted
to "fix":

  * save / load / bulk-load / update_label / unify_labels LOG their failures and return (hnsw_manager.py:69,111,132,
    198,224); a failed load leaves an empty index with ef=50 (hnsw_manager.py:69-76);
  * `find_similar_embeddings` ignores its `k`, searches min(50, count) and keeps ids with 1-d >= threshold - NON-strict,
    unlike the accept rule of the recogniser (hnsw_manager.py:227-244);
  * a full index only warns (hnsw_manager.py:135-143).

The body is organised differently from the reference: one byte-level `_BlobStore` does the optional encryption for
all three files, one `_relabel` transaction backs both the single rename and the group unification, one
`_decode_embedding` reads SQLite blobs for the bulk load and for `_get_embedding_from_db_id`.

Additive API (the B200 reasons to switch):
  * `max_elements` constructor argument instead of the hard-coded 100000 (hnsw_manager.py:29,43,62,71,136);
  * `query_batch(E, k)`: many queries, one launch chain;
  * `find_similar_all(threshold)`: the top-50 neighbour list of EVERY stored row in tiled GPU passes
    (fire_knn_search_rows) - the bulk form of the O(N) `find_similar_embeddings` calls that
    `FaceRecognition.shrink_db_ids` makes (modules/face_recognition.py:265-315);
  * `shrink_db_ids(cursor, conn, threshold)`: that whole clean-up job on top of it, same grouping decisions.
"""
from __future__ import annotations

import logging
import os
import pickle
import tempfile
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import hnswlib_compat as hnswlib

log = logging.getLogger()          # the reference logs through the root logger (logging.info(...))
SIMILAR_K = 50                     # hnsw_manager.py:236: the neighbourhood size is fixed, whatever `k` the caller passes


def _is_unknown(label: str) -> bool:
    return label.lower().startswith("unknown")


def _decode_embedding(blob, dim: Optional[int] = None):
    """SQLite BLOB -> (unit-norm float32 vector | None, reason).  Zero vectors and wrong sizes are unusable."""
    vec = np.frombuffer(blob, dtype=np.float32)
    if dim is not None and vec.shape[0] != dim:
        return None, "size"
    n = np.linalg.norm(vec)
    if n == 0:
        return None, "zero"
    return vec / n, ""


class _BlobStore:
    """Whole-file reads and writes, through the caller's encryptor when there is one (modules/encryption.py interface)."""

    def __init__(self, encryptor):
        self.encryptor = encryptor

    def get(self, path: str) -> bytes:
        if self.encryptor:
            return self.encryptor.read_and_decrypt(path)
        with open(path, "rb") as fh:
            return fh.read()

    def put(self, path: str, payload: bytes) -> None:
        if self.encryptor:
            self.encryptor.encrypt_and_write(path, payload)
            return
        with open(path, "wb") as fh:
            fh.write(payload)

    @staticmethod
    def via_tempfile(fn, payload: Optional[bytes] = None) -> bytes:
        """hnswlib-style indexes only speak file paths: run fn(path) on a scratch file, return what is in it afterwards."""
        fd, path = tempfile.mkstemp()
        try:
            with os.fdopen(fd, "wb") as fh:
                if payload is not None:
                    fh.write(payload)
            fn(path)
            with open(path, "rb") as fh:
                return fh.read()
        finally:
            os.remove(path)


class HNSWManager:
    def __init__(self, embedding_dim: int, hnsw_index_path: str, hnsw_labels_path: str,
                 hnsw_db_ids_path: str, encryptor, hnsw_ef_construction: int = 200, hnsw_m: int = 16,
                 max_elements: int = 100000):
        self.embedding_dim = embedding_dim
        self.hnsw_index_path, self.hnsw_labels_path, self.hnsw_db_ids_path = hnsw_index_path, hnsw_labels_path, hnsw_db_ids_path
        self.encryptor = encryptor
        self.max_elements = int(max_elements)
        self._store = _BlobStore(encryptor)
        self._graph = (int(hnsw_ef_construction), int(hnsw_m))
        self.hnsw_index = hnswlib.Index(space="cosine", dim=embedding_dim)
        self.hnsw_labels: List[str] = []
        self.hnsw_db_ids: List[int] = []
        self.hnsw_id_counter = 0
        if self._files_exist((hnsw_index_path, hnsw_labels_path, hnsw_db_ids_path)):
            self._load_hnswlib_index()
            log.info("Loaded existing HNSWlib index and mappings from disk.")
        else:
            self._fresh_index(ef=200, graph=self._graph)
            log.info("Initialized new HNSWlib index.")

    # ---- state -------------------------------------------------------------------------------------------------------
    def _fresh_index(self, ef: int, graph=(200, 16)) -> None:
        self.hnsw_index.init_index(max_elements=self.max_elements, ef_construction=graph[0], M=graph[1])
        self.hnsw_index.set_ef(ef)
        self.hnsw_labels, self.hnsw_db_ids, self.hnsw_id_counter = [], [], 0

    def _enrol(self, vectors: np.ndarray, labels: Sequence[str], db_ids: Sequence[int]) -> None:
        first = self.hnsw_id_counter
        self.hnsw_index.add_items(vectors, np.arange(first, first + len(labels), dtype=np.uint64))
        self.hnsw_labels.extend(labels)
        self.hnsw_db_ids.extend(db_ids)
        self.hnsw_id_counter = first + len(labels)

    def _files_exist(self, paths):
        return all(os.path.exists(p) for p in paths)

    # ---- persistence (hnsw_manager.py:36-112) --------------------------------------------------------------------------
    def _load_hnswlib_index(self):
        try:
            blob = self._store.get(self.hnsw_index_path)
            self._store.via_tempfile(lambda p: self.hnsw_index.load_index(p, max_elements=self.max_elements), blob)
            labels = pickle.loads(self._store.get(self.hnsw_labels_path))
            db_ids = pickle.loads(self._store.get(self.hnsw_db_ids_path))
            self.hnsw_labels, self.hnsw_db_ids, self.hnsw_id_counter = labels, db_ids, len(labels)
            log.info("Loaded HNSWlib index and mappings from disk.")
        except Exception as e:
            log.error(f"Error loading HNSWlib index: {e}")
            self._fresh_index(ef=50)                                   # hnsw_manager.py:71-72: ef 50 on this path, graph defaults
            log.info("Initialized a new HNSWlib index due to loading failure.")

    def save_hnswlib_index(self):
        try:
            files = ((self.hnsw_index_path, self._store.via_tempfile(self.hnsw_index.save_index)),
                     (self.hnsw_labels_path, pickle.dumps(self.hnsw_labels)),
                     (self.hnsw_db_ids_path, pickle.dumps(self.hnsw_db_ids)))
            for path, payload in files:
                self._store.put(path, payload)
            log.info("Saved HNSWlib index and mappings to disk.")
        except Exception as e:
            log.error(f"Error saving HNSWlib index: {e}")

    # ---- enrol (hnsw_manager.py:114-143) -------------------------------------------------------------------------------
    def load_embeddings_into_hnswlib(self, rows):
        """(db_id, label, float32 BLOB) rows out of SQLite.  Unusable rows are skipped with a warning; the usable ones are
        appended by ONE GPU call instead of one add_items per row."""
        try:
            vecs, labels, db_ids = [], [], []
            for db_id, label, blob in rows:
                vec, why = _decode_embedding(blob, self.embedding_dim)
                if vec is None:
                    log.warning(f"Embedding size mismatch for label '{label}'. Skipping." if why == "size"
                                else f"Zero vector found for label '{label}'. Skipping.")
                    continue
                vecs.append(vec); labels.append(label); db_ids.append(db_id)
            if vecs:
                self._enrol(np.stack(vecs).astype(np.float32), labels, db_ids)
            log.info("Loaded embeddings into HNSWlib index from SQLite database.")
        except Exception as e:
            log.error(f"Error loading embeddings into HNSWlib: {e}")

    def add_embedding(self, embedding: np.ndarray, label: str, db_id: int):
        if self.hnsw_id_counter >= self.max_elements:
            log.warning("HNSWlib index has reached its maximum capacity. Cannot add more embeddings.")
            return
        self._enrol(np.asarray(embedding, dtype=np.float32).reshape(1, -1), [label], [db_id])
        log.info(f"Added '{label}' to HNSWlib index with hnsw_id {self.hnsw_id_counter - 1}.")

    # ---- match (hnsw_manager.py:145-149, 227-244) ----------------------------------------------------------------------
    def query(self, embedding: np.ndarray, k=1):
        if self.hnsw_index.get_current_count() == 0:
            return None, None
        return self.hnsw_index.knn_query(embedding, k=k)

    def query_batch(self, embeddings: np.ndarray, k: int = 1):
        """[Q,D] -> (labels uint64 [Q,k], distances float32 [Q,k]) in one launch chain; (None, None) on an empty index."""
        return self.query(np.asarray(embeddings, dtype=np.float32).reshape(-1, self.embedding_dim), k=k)

    @staticmethod
    def _above(labels_row, dist_row, threshold: float) -> list:
        sims = 1 - np.asarray(dist_row)
        return [labels_row[j] for j in np.flatnonzero(sims >= threshold)]       # ascending distance order is kept

    def find_similar_embeddings(self, reference_embedding: np.ndarray, similarity_threshold: float, k: int = 50) -> list:
        n = self.hnsw_index.get_current_count()
        if n == 0:
            return []
        labels, distances = self.hnsw_index.knn_query(reference_embedding, k=min(SIMILAR_K, n))
        return self._above(labels[0], distances[0], similarity_threshold)

    def find_similar_all(self, similarity_threshold: float, tile: int = 8192) -> List[list]:
        """`find_similar_embeddings` of every stored row at once: result[i] is what the reference's loop gets for hnsw id i
        when it queries with that row's own (stored, normalised) vector.  GPU work: ceil(N / tile) top-50 searches whose
        queries are the stored rows themselves (fire_knn_search_rows)."""
        n = self.hnsw_index.get_current_count()
        out: List[list] = []
        for first in range(0, n, tile):
            labels, distances = self.hnsw_index.knn_query_rows(first, min(tile, n - first), k=min(SIMILAR_K, n))
            out.extend(self._above(labels[i], distances[i], similarity_threshold) for i in range(len(labels)))
        return out

    # ---- label maintenance (hnsw_manager.py:151-226) -------------------------------------------------------------------
    def _relabel(self, hnsw_ids: Iterable[int], new_label: str, db_cursor, db_conn) -> None:
        """One transaction: SQLite rows first, then the in-memory labels, then the on-disk copy."""
        ids = [int(h) for h in hnsw_ids]
        db_ids = [self.hnsw_db_ids[h] for h in ids]                   # a bad id fails here, before anything is written
        for db_id in db_ids:
            db_cursor.execute("UPDATE faces SET label = ? WHERE id = ?", (new_label, db_id))
        db_conn.commit()
        for hid in ids:
            self.hnsw_labels[hid] = new_label
        self.save_hnswlib_index()

    def _group_conflicts(self, hnsw_ids) -> bool:
        """More than one distinct KNOWN label inside the group: the reference refuses to merge such a group."""
        return len({self.hnsw_labels[h] for h in hnsw_ids if not _is_unknown(self.hnsw_labels[h])}) > 1

    def update_label(self, hnsw_id: int, new_label: str, db_cursor, db_conn, similarity_threshold: float = 0.7):
        try:
            if not 0 <= hnsw_id < len(self.hnsw_db_ids):
                log.error("Invalid hnsw_id for update_label.")
                return
            anchor = self._get_embedding_from_db_id(self.hnsw_db_ids[hnsw_id], db_cursor)
            group = self.find_similar_embeddings(anchor, similarity_threshold, k=SIMILAR_K) if anchor is not None else []
            if group and self._group_conflicts(group):
                log.warning("Conflicting known labels found. Not unifying this group.")
                group = []
            if group:
                self.unify_labels(group, new_label, db_cursor, db_conn)
            else:                                  # no stored vector, no neighbour, or a conflict: only the requested row
                self._rename_single_entry(hnsw_id, new_label, db_cursor, db_conn)
        except Exception as e:
            log.error(f"Error updating label: {e}")

    def _rename_single_entry(self, hnsw_id, new_label, db_cursor, db_conn):
        self._relabel([hnsw_id], new_label, db_cursor, db_conn)
        log.info(f"Updated label for hnsw_id {hnsw_id} (db_id {self.hnsw_db_ids[hnsw_id]}) to '{new_label}'.")

    def unify_labels(self, hnsw_ids: list, new_label: str, db_cursor, db_conn):
        try:
            self._relabel(hnsw_ids, new_label, db_cursor, db_conn)
            log.info(f"Unified {len(hnsw_ids)} embeddings under label '{new_label}'.")
        except Exception as e:
            log.error(f"Error unifying labels: {e}")

    def _get_embedding_from_db_id(self, db_id: int, db_cursor):
        try:
            db_cursor.execute("SELECT embedding FROM faces WHERE id=?", (db_id,))
            row = db_cursor.fetchone()
            if row:
                vec, why = _decode_embedding(row[0])
                return vec if vec is not None else np.frombuffer(row[0], dtype=np.float32)     # a zero vector comes back as it is
        except Exception as e:
            log.error(f"Error retrieving embedding from DB: {e}")
        return None

    def shrink_db_ids(self, db_cursor, db_conn, similarity_threshold: float = 0.75) -> int:
        """`FaceRecognition.shrink_db_ids` (modules/face_recognition.py:265-315) as a bulk job: every row's neighbour list
        comes from `find_similar_all` (a handful of GPU passes) instead of one SQLite read + one index query per row; the
        walk over the ids - skip what an earlier group already covered, refuse groups with two different known labels,
        otherwise merge under the known label or the anchor's own - is the reference's, decision for decision.
        Rows whose SQLite record is missing are skipped like there.  Returns the number of merges.
        (The reference re-reads each vector from SQLite and normalises it with numpy before hnswlib normalises it again;
        the stored row differs from that by one rounding, ~1e-7 in distance - inside the 1e-5 tie window of the parity bar.)"""
        neighbours = self.find_similar_all(similarity_threshold)
        done, merges = set(), 0
        for hid, group in enumerate(neighbours):
            if hid in done:
                continue
            if self._get_embedding_from_db_id(self.hnsw_db_ids[hid], db_cursor) is None:
                continue
            group = [int(g) for g in group]
            if len(group) <= 1:
                done.add(hid)
                continue
            done.update(group)
            if self._group_conflicts(group):
                continue
            known = [self.hnsw_labels[g] for g in group if not _is_unknown(self.hnsw_labels[g])]
            self.unify_labels(group, known[0] if known else self.hnsw_labels[hid], db_cursor, db_conn)
            merges += 1
        log.info(f"DB ID shrinking completed with {merges} unification operations.")
        return merges
