#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "block8 or config2 or repeated or golden" > gpurun_out/s31_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s31_pytest.log
tail -6 gpurun_out/s31_pytest.log
timeout 300 python tools/trace_block8.py 256 2> gpurun_out/s31_block8_timeline.txt; tail -7 gpurun_out/s31_block8_timeline.txt | cut -c1-250
for cfg in "FIRE_B200_B8_MC=0" "FIRE_B200_B8_MC=1" "FIRE_B200_B8_MC=0" "FIRE_B200_B8_MC=1"; do
  env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$cfg]', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
done
