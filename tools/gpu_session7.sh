#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q -x --timeout=600 > gpurun_out/s7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -5 gpurun_out/s7_pytest.log
for pr in 0 1; do
  for q in 1 32 128 256 1024 4096; do
    FIRE_B200_KNN_PAIR=$pr timeout 200 python tools/knn_probe.py 1000000 $q 20 >> gpurun_out/s7_knn.txt 2>&1
  done
done
cat gpurun_out/s7_knn.txt
