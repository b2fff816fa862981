"""fire_b200: B200-native (sm_100a) identification hot path for FIRE.

Importable as ``fire_b200`` (see /fire_b200/__init__.py, a path alias: this directory's name
is fixed by the build contract and is not a valid Python identifier).

Public surfaces (drop-in for the reference):
  fire_b200.encoder.Encoder            <- modules/encoder.py:9-27
  fire_b200.facenet_gpu.FaceNetClient  <- facenet_gpu.py:84-129
  fire_b200.hnsw_manager.HNSWManager   <- modules/hnsw_manager.py:11-262
  fire_b200.preprocess                 <- processing/preprocess.py:10-145 (+ crop_resize_normalize)
Everything computes through libfire_b200.so (include/fire_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
