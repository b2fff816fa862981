"""recall@10 of "the reference's HNSW index" against the exact result (north star: "recall@10 is reported against
the reference's HNSW index").  CPU only: oracle/hnsw_oracle.c (hnswlib Index restated: M 16, ef_construction 200,
seed 100, modules/hnsw_manager.py:29-30) vs oracle BFIndex (exact).  fire_b200's kNN returns the exact BFIndex ids
(GPU parity tests), so its recall@10 against the exact result is 1.0 by construction.

    python tools/recall_report.py > profiles/rNN_recall_vs_hnsw.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402

from oracle import native   # noqa: E402


def face_like(n_id, per_id, D, rng, noise=0.35):
    """unit-norm identity centres + per-sample noise: the clustered structure face embeddings have"""
    c = rng.standard_normal((n_id, D)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    x = np.repeat(c, per_id, axis=0) + noise / np.sqrt(D) * rng.standard_normal((n_id * per_id, D)).astype(np.float32) * np.sqrt(D) * 0.2
    return x.astype(np.float32), c


def report(name, gallery, queries, k=10):
    D = gallery.shape[1]
    t = time.time()
    h = native.HnswOracle(D, max_elements=max(100000, len(gallery)))
    h.add_items(gallery)
    tb = time.time() - t
    bf = native.BFIndexOracle(D)
    bf.add_items(gallery)
    t = time.time()
    el, _ = bf.knn_query(queries, k)
    te = time.time() - t
    print(f"## {name}: gallery {gallery.shape}, {len(queries)} queries, k={k}  (HNSW build {tb:.1f} s, exact scan {te:.2f} s on the CPU)")
    for ef, why in ((200, "set_ef(200), a freshly built index (hnsw_manager.py:30)"),
                    (50, "ef 50, the load-failure path (hnsw_manager.py:72)"),
                    (10, "ef 10, hnswlib's default after load_index - the reference forgets set_ef (hnsw_manager.py:43,62)")):
        h.set_ef(ef)
        t = time.time()
        l, _ = h.knn_query(queries, k)
        tq = time.time() - t
        rec = float(np.mean([len(set(l[i].tolist()) & set(el[i].tolist())) / k for i in range(len(queries))]))
        top1 = float(np.mean(l[:, 0] == el[:, 0]))
        print(f"  reference HNSW  {why:95s} recall@{k} {rec:.4f}   top-1 agreement {top1:.4f}   {len(queries) / tq:8.0f} QPS (1 CPU thread)")
    print(f"  fire_b200 (exact brute force on the B200, ids == BFIndex)                                                       recall@{k} 1.0000   top-1 agreement 1.0000")


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    print("# recall@10 of the reference's approximate index vs the exact result; CPU restatement of hnswlib (parity unpinned, see oracle/hnsw_oracle.c)")
    g = rng.standard_normal((10_000, 128)).astype(np.float32)
    report("configs[0]-style gallery, isotropic Gaussian rows (worst case for a graph index)", g, rng.standard_normal((500, 128)).astype(np.float32))
    gal, centres = face_like(5_000, 8, 512, rng)
    q = centres[rng.integers(0, len(centres), 500)] + 0.07 * rng.standard_normal((500, 512)).astype(np.float32)
    report("face-like gallery: 5000 identities x 8 enrolled embeddings, 512-d", gal, q.astype(np.float32))
