"""Make the reference's own callers (modules/face_recognition.py, main.py) run on the B200 path.

    import fire_b200.dropin as dropin
    dropin.install()            # before `from modules.face_recognition import FaceRecognition`

Two levels:
  level="engines" (default)  registers `facenet_gpu` and `hnswlib` in sys.modules, so the reference's UNMODIFIED
                             modules/encoder.py and modules/hnsw_manager.py run on top of the sm_100a engines
                             (cv2.resize stays where the reference calls it);
  level="modules"            additionally replaces `modules.encoder` and `modules.hnsw_manager` with the
                             fire_b200 versions, which also move preprocess_for_encoder to the GPU and lift the
                             hard-coded 100000-row capacity.
"""
from __future__ import annotations

import sys
import types


def install(level: str = "engines") -> None:
    if level not in ("engines", "modules"):
        raise ValueError("level must be 'engines' or 'modules'")
    from . import encoder, facenet_gpu, hnsw_manager, hnswlib_compat
    sys.modules["facenet_gpu"] = facenet_gpu
    sys.modules["hnswlib"] = hnswlib_compat
    if level == "modules":
        pkg = sys.modules.get("modules")
        if pkg is None:
            pkg = types.ModuleType("modules")
            pkg.__path__ = []
            sys.modules["modules"] = pkg
        sys.modules["modules.encoder"] = encoder
        sys.modules["modules.hnsw_manager"] = hnsw_manager
        pkg.encoder, pkg.hnsw_manager = encoder, hnsw_manager


def uninstall() -> None:
    for name in ("facenet_gpu", "hnswlib", "modules.encoder", "modules.hnsw_manager"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith("fire_b200"):
            del sys.modules[name]
