#!/bin/bash
mkdir -p gpurun_out
for cfg in "FIRE_B200_B35_COOP=0" "FIRE_B200_B35_COOP=1" "FIRE_B200_B35_COOP=0" "FIRE_B200_B35_COOP=1"; do
  env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$cfg]', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['min_cos_vs_fp32_oracle'])"
done 2>&1 | tee gpurun_out/s40_coop.txt
FIRE_B200_B35_COOP=1 timeout 600 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "block35 or repeated or config2" 2>&1 | tail -3
