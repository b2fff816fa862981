#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py tests/test_gpu_dropin.py -m gpu -q -x --timeout=300 > gpurun_out/s18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s18_pytest.log
tail -4 gpurun_out/s18_pytest.log
for cfg in "FIRE_B200_SIDE_POOLS=0" "FIRE_B200_SIDE_POOLS=1" "FIRE_B200_SIDE_POOLS=0" "FIRE_B200_SIDE_POOLS=1"; do
  env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$cfg]', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
done
