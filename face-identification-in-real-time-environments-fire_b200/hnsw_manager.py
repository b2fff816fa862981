"""Drop-in for the reference's ``modules/hnsw_manager.py`` (modules/hnsw_manager.py:11-262).

Same constructor, attributes (`hnsw_index`, `hnsw_labels`, `hnsw_db_ids`, `hnsw_id_counter`) and
methods with the reference's semantics, including the parts a rewrite would be tempted to "fix":
errors in save/load/update are logged, not raised (hnsw_manager.py:69,111,132,198,224);
`find_similar_embeddings` ignores its `k` and searches min(50, count) with a NON-strict `>=`
(hnsw_manager.py:227-244); a failed load starts an empty index with ef=50 (hnsw_manager.py:69-76).

Differences, all additive:
  * the index is the exact B200 cosine search (fire_b200.hnswlib_compat.Index), so results equal
    hnswlib.BFIndex rather than an approximate graph walk;
  * capacity is a constructor argument (`max_elements`, default 100000 - the value hard-coded at
    hnsw_manager.py:29,43,62,71,136);
  * `query_batch` answers many queries in one GPU call.
"""
from __future__ import annotations

import logging
import os
import pickle
import tempfile

import numpy as np

from . import hnswlib_compat as hnswlib


class HNSWManager:
    def __init__(self, embedding_dim: int, hnsw_index_path: str, hnsw_labels_path: str,
                 hnsw_db_ids_path: str, encryptor, hnsw_ef_construction: int = 200, hnsw_m: int = 16,
                 max_elements: int = 100000):
        self.embedding_dim = embedding_dim
        self.hnsw_index_path = hnsw_index_path
        self.hnsw_labels_path = hnsw_labels_path
        self.hnsw_db_ids_path = hnsw_db_ids_path
        self.encryptor = encryptor
        self.max_elements = int(max_elements)

        self.hnsw_index = hnswlib.Index(space='cosine', dim=self.embedding_dim)
        self.hnsw_labels = []
        self.hnsw_db_ids = []
        self.hnsw_id_counter = 0

        if self._files_exist([self.hnsw_index_path, self.hnsw_labels_path, self.hnsw_db_ids_path]):
            self._load_hnswlib_index()
            logging.info("Loaded existing HNSWlib index and mappings from disk.")
        else:
            self.hnsw_index.init_index(max_elements=self.max_elements, ef_construction=hnsw_ef_construction, M=hnsw_m)
            self.hnsw_index.set_ef(200)
            logging.info("Initialized new HNSWlib index.")

    # ---- persistence ---------------------------------------------------------------------------
    def _files_exist(self, paths):
        return all(os.path.exists(path) for path in paths)

    def _read_file(self, path: str) -> bytes:
        if self.encryptor:
            return self.encryptor.read_and_decrypt(path)
        with open(path, 'rb') as f:
            return f.read()

    def _write_file(self, path: str, data: bytes):
        if self.encryptor:
            self.encryptor.encrypt_and_write(path, data)
        else:
            with open(path, 'wb') as f:
                f.write(data)

    def _load_hnswlib_index(self):
        try:
            index_data = self._read_file(self.hnsw_index_path)
            with tempfile.NamedTemporaryFile(delete=False) as tmp_index:
                tmp_index.write(index_data)
                tmp_index_path = tmp_index.name
            try:
                self.hnsw_index.load_index(tmp_index_path, max_elements=self.max_elements)
            finally:
                os.remove(tmp_index_path)
            self.hnsw_labels = pickle.loads(self._read_file(self.hnsw_labels_path))
            self.hnsw_db_ids = pickle.loads(self._read_file(self.hnsw_db_ids_path))
            self.hnsw_id_counter = len(self.hnsw_labels)
            logging.info("Loaded HNSWlib index and mappings from disk.")
        except Exception as e:
            logging.error(f"Error loading HNSWlib index: {e}")
            self.hnsw_index.init_index(max_elements=self.max_elements, ef_construction=200, M=16)
            self.hnsw_index.set_ef(50)
            self.hnsw_labels = []
            self.hnsw_db_ids = []
            self.hnsw_id_counter = 0
            logging.info("Initialized a new HNSWlib index due to loading failure.")

    def save_hnswlib_index(self):
        try:
            with tempfile.NamedTemporaryFile(delete=False) as tmp_index:
                tmp_index_path = tmp_index.name
            try:
                self.hnsw_index.save_index(tmp_index_path)
                with open(tmp_index_path, 'rb') as f:
                    index_data = f.read()
            finally:
                os.remove(tmp_index_path)
            self._write_file(self.hnsw_index_path, index_data)
            self._write_file(self.hnsw_labels_path, pickle.dumps(self.hnsw_labels))
            self._write_file(self.hnsw_db_ids_path, pickle.dumps(self.hnsw_db_ids))
            logging.info("Saved HNSWlib index and mappings to disk.")
        except Exception as e:
            logging.error(f"Error saving HNSWlib index: {e}")

    # ---- enrol -----------------------------------------------------------------------------------
    def load_embeddings_into_hnswlib(self, rows):
        """rows of (db_id, label, float32 blob) from SQLite (hnsw_manager.py:114-133); appended in one GPU call."""
        try:
            keep, labels, db_ids = [], [], []
            for db_id, label, embedding_blob in rows:
                embedding = np.frombuffer(embedding_blob, dtype=np.float32)
                if embedding.shape[0] != self.embedding_dim:
                    logging.warning(f"Embedding size mismatch for label '{label}'. Skipping.")
                    continue
                norm = np.linalg.norm(embedding)
                if norm == 0:
                    logging.warning(f"Zero vector found for label '{label}'. Skipping.")
                    continue
                keep.append(embedding / norm)
                labels.append(label)
                db_ids.append(db_id)
            if keep:
                ids = np.arange(self.hnsw_id_counter, self.hnsw_id_counter + len(keep), dtype=np.uint64)
                self.hnsw_index.add_items(np.stack(keep).astype(np.float32), ids)
                self.hnsw_labels.extend(labels)
                self.hnsw_db_ids.extend(db_ids)
                self.hnsw_id_counter += len(keep)
            logging.info("Loaded embeddings into HNSWlib index from SQLite database.")
        except Exception as e:
            logging.error(f"Error loading embeddings into HNSWlib: {e}")

    def add_embedding(self, embedding: np.ndarray, label: str, db_id: int):
        if self.hnsw_id_counter < self.max_elements:
            self.hnsw_index.add_items(embedding, self.hnsw_id_counter)
            self.hnsw_labels.append(label)
            self.hnsw_db_ids.append(db_id)
            logging.info(f"Added '{label}' to HNSWlib index with hnsw_id {self.hnsw_id_counter}.")
            self.hnsw_id_counter += 1
        else:
            logging.warning("HNSWlib index has reached its maximum capacity. Cannot add more embeddings.")

    # ---- match -----------------------------------------------------------------------------------
    def query(self, embedding: np.ndarray, k=1):
        if self.hnsw_index.get_current_count() > 0:
            labels, distances = self.hnsw_index.knn_query(embedding, k=k)
            return labels, distances
        return None, None

    def query_batch(self, embeddings: np.ndarray, k: int = 1):
        """[Q,D] -> (labels uint64 [Q,k], distances float32 [Q,k]) in one launch chain (additive API)."""
        return self.query(np.asarray(embeddings, dtype=np.float32).reshape(-1, self.embedding_dim), k=k)

    def find_similar_embeddings(self, reference_embedding: np.ndarray, similarity_threshold: float, k: int = 50) -> list:
        count = self.hnsw_index.get_current_count()
        if count == 0:
            return []
        k_search = min(50, count)                       # the reference ignores `k` (hnsw_manager.py:236)
        labels, distances = self.hnsw_index.knn_query(reference_embedding, k=k_search)
        similar_ids = []
        for i in range(len(labels[0])):
            sim = 1 - distances[0][i]
            if sim >= similarity_threshold:             # non-strict, unlike the accept rule elsewhere
                similar_ids.append(labels[0][i])
        return similar_ids

    # ---- label maintenance -------------------------------------------------------------------------
    def update_label(self, hnsw_id: int, new_label: str, db_cursor, db_conn, similarity_threshold: float = 0.7):
        try:
            if hnsw_id < 0 or hnsw_id >= len(self.hnsw_db_ids):
                logging.error("Invalid hnsw_id for update_label.")
                return
            reference_embedding = self._get_embedding_from_db_id(self.hnsw_db_ids[hnsw_id], db_cursor)
            if reference_embedding is None:
                self._rename_single_entry(hnsw_id, new_label, db_cursor, db_conn)
                return
            similar_ids = self.find_similar_embeddings(reference_embedding, similarity_threshold, k=50)
            if not similar_ids:
                self._rename_single_entry(hnsw_id, new_label, db_cursor, db_conn)
                return
            known = {self.hnsw_labels[sid] for sid in similar_ids
                     if not self.hnsw_labels[sid].lower().startswith("unknown")}
            if len(known) > 1:
                logging.warning("Conflicting known labels found. Not unifying this group.")
                self._rename_single_entry(hnsw_id, new_label, db_cursor, db_conn)
                return
            self.unify_labels(similar_ids, new_label, db_cursor, db_conn)
        except Exception as e:
            logging.error(f"Error updating label: {e}")

    def _rename_single_entry(self, hnsw_id, new_label, db_cursor, db_conn):
        db_id = self.hnsw_db_ids[hnsw_id]
        db_cursor.execute('UPDATE faces SET label = ? WHERE id = ?', (new_label, db_id))
        db_conn.commit()
        self.hnsw_labels[hnsw_id] = new_label
        logging.info(f"Updated label for hnsw_id {hnsw_id} (db_id {db_id}) to '{new_label}'.")
        self.save_hnswlib_index()

    def unify_labels(self, hnsw_ids: list, new_label: str, db_cursor, db_conn):
        try:
            for hid in hnsw_ids:
                db_cursor.execute('UPDATE faces SET label = ? WHERE id = ?', (new_label, self.hnsw_db_ids[hid]))
            db_conn.commit()
            for hid in hnsw_ids:
                self.hnsw_labels[hid] = new_label
            logging.info(f"Unified {len(hnsw_ids)} embeddings under label '{new_label}'.")
            self.save_hnswlib_index()
        except Exception as e:
            logging.error(f"Error unifying labels: {e}")

    def _get_embedding_from_db_id(self, db_id: int, db_cursor):
        try:
            db_cursor.execute('SELECT embedding FROM faces WHERE id=?', (db_id,))
            row = db_cursor.fetchone()
            if row:
                embedding = np.frombuffer(row[0], dtype=np.float32)
                norm = np.linalg.norm(embedding)
                if norm > 0:
                    embedding = embedding / norm
                return embedding
        except Exception as e:
            logging.error(f"Error retrieving embedding from DB: {e}")
        return None
