#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
tail -15 gpurun_out/s5_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s5_bench.err
for sh in 0 1; do
  for q in 256 1024 4096; do
    FIRE_B200_KNN_PAIR=$sh timeout 200 python tools/knn_probe.py 1000000 $q 20 >> gpurun_out/s5_knn_share.txt 2>&1
  done
done
cat gpurun_out/s5_knn_share.txt
