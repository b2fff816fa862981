"""an index file written by hnswlib 0.8.0 `save_index` loads into the drop-in.  hnswlib is not
installable here, so the file is produced by a restatement of HierarchicalNSW::saveIndex (tests only; PARITY UNPINNED
against a real hnswlib file) from a synthetic gallery with dummy graph links."""
import struct

import numpy as np
import pytest

import fakes


def write_hnswlib_file(path, rows, labels, M=16, max_elements=100000, deleted=()):
    n, dim = rows.shape
    maxM0 = 2 * M
    data_off = maxM0 * 4 + 4
    label_off = data_off + dim * 4
    per_el = label_off + 8
    rng = np.random.default_rng(0)
    with open(path, "wb") as f:
        f.write(struct.pack("<6QiI3QdQ", 0, max_elements, n, per_el, label_off, data_off, 1, 0, M, maxM0, M, 1.0 / np.log(M), 200))
        for i in range(n):
            nl = int(rng.integers(0, maxM0 + 1))
            head = nl | ((1 << 16) if i in deleted else 0)
            links = np.zeros(maxM0, dtype="<u4")
            links[:nl] = rng.integers(0, n, nl)
            f.write(struct.pack("<I", head) + links.tobytes() + rows[i].astype("<f4").tobytes() + struct.pack("<Q", int(labels[i])))
        for i in range(n):                                   # upper levels: element 0 lives on level 1
            if i == 0:
                f.write(struct.pack("<I", M * 4 + 4) + bytes(M * 4 + 4))
            else:
                f.write(struct.pack("<I", 0))


def test_hnswlib_binary_loads_into_the_dropin(tmp_path, monkeypatch, oracle_native):
    from fire_b200 import hnswlib_compat
    monkeypatch.setattr(hnswlib_compat._engine, "KnnIndex", fakes.FakeKnnIndex)
    rng = np.random.default_rng(1)
    rows = oracle_native.normalize(rng.standard_normal((300, 128)).astype(np.float32))
    path = str(tmp_path / "hnsw_index_yunet_128.bin")
    write_hnswlib_file(path, rows, np.arange(300), deleted={7})
    idx = hnswlib_compat.Index(space="cosine", dim=128)
    idx.load_index(path, max_elements=100000)                # hnsw_manager.py:43,62
    assert idx.get_current_count() == 299 and idx.get_max_elements() == 100000 and idx.ef == 10
    q = rng.standard_normal((20, 128)).astype(np.float32)
    lab, dist = idx.knn_query(q, k=5)
    ora = oracle_native.BFIndexOracle(128)
    keep = np.r_[0:7, 8:300]
    ora.add_items(rows[keep])
    ol, od = ora.knn_query(q, 5)
    assert np.array_equal(lab, keep[ol.astype(np.int64)].astype(np.uint64)) and np.abs(dist - od).max() < 5e-6
    # wrong dimension / garbage are rejected like hnswlib rejects a corrupt file
    with pytest.raises(RuntimeError):
        hnswlib_compat.Index(space="cosine", dim=512).load_index(path)
    (tmp_path / "junk.bin").write_bytes(b"\x01" * 500)
    with pytest.raises(RuntimeError):
        hnswlib_compat.Index(space="cosine", dim=128).load_index(str(tmp_path / "junk.bin"))
    # what the drop-in saves is an hnswlib 0.8.0 image too (a roll-back to the reference must not lose the gallery):
    # same header arithmetic, every element on level 0, neighbour lists = the exact 2M nearest rows, no self links
    idx.save_index(str(tmp_path / "own.bin"))
    blob = (tmp_path / "own.bin").read_bytes()
    hdr = hnswlib_compat._HNSW_HEADER.unpack_from(blob, 0)
    n, per_el, label_off, data_off = hdr[2], hdr[3], hdr[4], hdr[5]
    assert (hdr[0], n, hdr[8], hdr[9], hdr[10]) == (0, 299, 16, 32, 16) and hdr[6] == 0 and hdr[7] == 0
    assert len(blob) == hnswlib_compat._HNSW_HEADER.size + n * per_el + 4 * n            # + one empty upper-level list per element
    rec = np.frombuffer(blob, np.uint8, n * per_el, hnswlib_compat._HNSW_HEADER.size).reshape(n, per_el)
    head = np.ascontiguousarray(rec[:, :data_off]).view("<u4").reshape(n, 33)
    assert np.all(head[:, 0] == 32) and np.all(head[:, 1:] < n) and not np.any(head[:, 1:] == np.arange(n)[:, None])
    stored = np.ascontiguousarray(rec[:, data_off:label_off]).view("<f4").reshape(n, 128)
    want_nb = np.argsort(-(stored @ stored.T) + 2 * np.eye(n), axis=1)[:, :32]           # brute-force neighbours, self pushed last
    assert np.mean([len(set(head[i, 1:]) & set(want_nb[i])) for i in range(n)]) > 31.9
    again = hnswlib_compat.Index(space="cosine", dim=128)
    again.load_index(str(tmp_path / "own.bin"))
    assert again.get_current_count() == 299 and np.array_equal(again.knn_query(q, k=5)[0], lab)
    # files written by round 1 of this package (private FIREKNN1 layout) still load
    old = tmp_path / "old.bin"
    old.write_bytes(b"FIREKNN1" + struct.pack("<qqqq", 128, 3, 1000, 10) + np.arange(3, dtype="<u8").tobytes() + rows[:3].astype("<f4").tobytes())
    legacy = hnswlib_compat.Index(space="cosine", dim=128)
    legacy.load_index(str(old))
    assert legacy.get_current_count() == 3 and legacy.knn_query(rows[1], k=1)[0][0][0] == 1
