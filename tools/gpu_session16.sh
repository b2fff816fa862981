#!/bin/bash
mkdir -p gpurun_out
for cfg in "" "FIRE_B200_STRIP_GROUP=1" "FIRE_B200_STRIP_ACC=4" "FIRE_B200_STRIP_ACC=2"; do
  echo "== [$cfg]"
  env $cfg timeout 300 python tools/profile_ops.py 256 512 2>&1 | sed -n 3,5p
done
for op in 1 2; do
  FIRE_B200_TRACE_OP=$op timeout 300 python tools/profile_ops.py 256 512 2>&1 >/dev/null | grep -A7 "trace op"
done
timeout 600 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "pair_stem or strip or config2" 2>&1 | tail -3
