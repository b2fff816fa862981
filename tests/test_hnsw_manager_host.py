"""Host-side logic of the HNSWManager drop-in (labels, db ids, persistence, thresholds, label maintenance) against
the reference semantics (modules/hnsw_manager.py), with the GPU index replaced by an oracle-backed double."""
import sqlite3

import numpy as np
import pytest

import fakes


@pytest.fixture()
def mgr_factory(tmp_path, monkeypatch, oracle_native):
    from fire_b200 import hnsw_manager, hnswlib_compat
    monkeypatch.setattr(hnswlib_compat._engine, "KnnIndex", fakes.FakeKnnIndex)

    def make(dim=16, encryptor=None, **kw):
        return hnsw_manager.HNSWManager(dim, str(tmp_path / "i.bin"), str(tmp_path / "l.pkl"), str(tmp_path / "d.pkl"), encryptor, **kw)
    return make


def _unit(rng, d=16):
    v = rng.standard_normal(d).astype(np.float32)
    return v / np.linalg.norm(v)


def test_add_query_and_capacity(mgr_factory):
    rng = np.random.default_rng(0)
    m = mgr_factory(max_elements=3)
    assert m.query(_unit(rng)) == (None, None)                                   # hnsw_manager.py:145-149 empty index
    vs = [_unit(rng) for _ in range(4)]
    for i, v in enumerate(vs):
        m.add_embedding(v * (i + 1.0), f"p{i}", 100 + i)                         # un-normalised input is normalised inside
    assert m.hnsw_id_counter == 3 and m.hnsw_labels == ["p0", "p1", "p2"] and m.hnsw_db_ids == [100, 101, 102]
    assert m.hnsw_index.get_current_count() == 3                                 # 4th add: capacity warning, not an exception
    labels, dist = m.query(vs[1], k=1)
    assert labels.dtype == np.uint64 and labels.shape == (1, 1) and labels[0][0] == 1 and abs(dist[0][0]) < 1e-6
    assert labels.size > 0 and m.hnsw_labels[labels[0][0]] == "p1"               # how face_recognition.py:460-465 uses it


def test_find_similar_is_nonstrict_and_ignores_k(mgr_factory):
    rng = np.random.default_rng(1)
    m = mgr_factory()
    base = _unit(rng)
    rows = []
    for c in [1.0, 0.9, 0.7, 0.5]:
        r = _unit(rng); r -= r.dot(base) * base; r /= np.linalg.norm(r)
        rows.append((c * base + np.sqrt(1 - c * c) * r).astype(np.float32))
    for i, r in enumerate(rows):
        m.add_embedding(r, f"x{i}", i)
    _, d = m.query(base, k=4)
    thr = float(1 - d[0][2])                                                     # exactly the similarity of the third row
    assert m.find_similar_embeddings(base, thr, k=1) == [0, 1, 2]               # >= (non-strict), and k is ignored (min(50, count))
    assert m.find_similar_embeddings(base, np.nextafter(np.float32(thr), np.float32(2))) == [0, 1]


def test_save_load_roundtrip_plain_and_encrypted(mgr_factory):
    class XorEncryptor:                                                          # same interface as modules/encryption.py
        def encrypt_and_write(self, path, data):
            open(path, "wb").write(bytes(b ^ 0x5A for b in data))

        def read_and_decrypt(self, path):
            return bytes(b ^ 0x5A for b in open(path, "rb").read())

    rng = np.random.default_rng(2)
    for enc in (None, XorEncryptor()):
        m = mgr_factory(encryptor=enc)
        m.hnsw_index.init_index(max_elements=100000, ef_construction=200, M=16)
        m.hnsw_labels, m.hnsw_db_ids, m.hnsw_id_counter = [], [], 0
        vs = [_unit(rng) for _ in range(5)]
        for i, v in enumerate(vs):
            m.add_embedding(v, f"n{i}", 10 + i)
        m.save_hnswlib_index()
        m2 = mgr_factory(encryptor=enc)                                          # constructor finds the three files and loads them
        assert m2.hnsw_labels == m.hnsw_labels and m2.hnsw_db_ids == m.hnsw_db_ids and m2.hnsw_id_counter == 5
        l, d = m2.query(vs[3], k=1)
        assert l[0][0] == 3 and abs(d[0][0]) < 1e-6
        assert m2.hnsw_index.ef == 10                                            # hnswlib does not persist ef; the reference never re-sets it


def test_corrupt_index_starts_empty_with_ef50(mgr_factory, tmp_path):
    for n in ("i.bin", "l.pkl", "d.pkl"):
        (tmp_path / n).write_bytes(b"garbage")
    m = mgr_factory()
    assert m.hnsw_index.get_current_count() == 0 and m.hnsw_labels == [] and m.hnsw_index.ef == 50   # hnsw_manager.py:69-76


def test_bulk_load_skips_bad_rows(mgr_factory):
    rng = np.random.default_rng(3)
    m = mgr_factory()
    good = [_unit(rng) * 3 for _ in range(3)]
    rows = [(1, "a", good[0].tobytes()), (2, "bad-dim", np.ones(5, np.float32).tobytes()),
            (3, "zero", np.zeros(16, np.float32).tobytes()), (4, "b", good[1].tobytes()), (5, "c", good[2].tobytes())]
    m.load_embeddings_into_hnswlib(rows)                                         # hnsw_manager.py:114-133
    assert m.hnsw_labels == ["a", "b", "c"] and m.hnsw_db_ids == [1, 4, 5] and m.hnsw_id_counter == 3
    l, _ = m.query(good[1], k=1)
    assert l[0][0] == 1


def test_update_label_unifies_unknowns_but_not_conflicting_known_labels(mgr_factory):
    rng = np.random.default_rng(4)
    m = mgr_factory()
    conn = sqlite3.connect(":memory:")
    cur = conn.cursor()
    cur.execute("CREATE TABLE faces (id INTEGER PRIMARY KEY AUTOINCREMENT, label TEXT, embedding BLOB)")   # database.py:53-59
    base = _unit(rng)

    def near(c):
        r = _unit(rng); r -= r.dot(base) * base; r /= np.linalg.norm(r)
        return (c * base + np.sqrt(1 - c * c) * r).astype(np.float32)

    def enrol(vec, label):
        cur.execute("INSERT INTO faces (label, embedding) VALUES (?, ?)", (label, vec.tobytes()))
        m.add_embedding(vec, label, cur.lastrowid)

    enrol(near(1.0), "Unknown_aaaa"); enrol(near(0.95), "Unknown_bbbb"); enrol(near(0.2), "Unknown_cccc")
    m.update_label(0, "alice", cur, conn, similarity_threshold=0.7)              # hnsw_manager.py:151-199
    assert m.hnsw_labels == ["alice", "alice", "Unknown_cccc"]
    assert [r[0] for r in cur.execute("SELECT label FROM faces ORDER BY id")] == ["alice", "alice", "Unknown_cccc"]
    enrol(near(0.93), "bob")                                                     # a second KNOWN label inside the cluster
    m.update_label(1, "carol", cur, conn, similarity_threshold=0.7)              # conflict -> only the requested row is renamed
    assert m.hnsw_labels == ["alice", "carol", "Unknown_cccc", "bob"]
    m.update_label(99, "nobody", cur, conn)                                      # invalid id: logged, nothing raised
