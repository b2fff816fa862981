#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "pool_conv or config2 or golden or config1" > gpurun_out/s33_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s33_pytest.log
tail -12 gpurun_out/s33_pytest.log
timeout 300 python tools/profile_ops.py 256 512 2>&1 | sed -n 1,9p
for cfg in "FIRE_B200_FUSE_POOL=0" "FIRE_B200_FUSE_POOL=1" "FIRE_B200_FUSE_POOL=0" "FIRE_B200_FUSE_POOL=1"; do
  env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$cfg]', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
done
