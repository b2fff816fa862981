"""Build recipe for libfire_b200.so (hand-written sm_100a CUDA behind the C ABI of include/fire_b200.h).

    python -m fire_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE (next to this file) so
that it travels to the GPU box with the repo snapshot; it links the static CUDA runtime only and
resolves cuTensorMapEncodeTiled at run time, so it also loads (but cannot compute) on a host
without a driver.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "_obj")
LIB_PATH = os.path.join(PKG_DIR, "libfire_b200.so")
SOURCES = ["fire_api.cu", "knn.cu", "preprocess.cu", "roi_upload.cu", "facenet_engine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false" if False else "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newest_source_mtime() -> float:
    t = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    t = max(t, os.path.getmtime(os.path.join(os.path.dirname(PKG_DIR), "include", "fire_b200.h")))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
