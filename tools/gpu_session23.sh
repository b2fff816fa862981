#!/bin/bash
# block35_fused: balanced task order (chains wander over CTAs through per-(block, image) flags): parity, timeline, A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 > gpurun_out/s23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s23_pytest.log
tail -6 gpurun_out/s23_pytest.log
timeout 300 python tools/trace_block35.py 256 2> gpurun_out/s23_block35_timeline.txt; tail -12 gpurun_out/s23_block35_timeline.txt | cut -c1-220
for cfg in "FIRE_B200_B35_BALANCE=0" "FIRE_B200_B35_BALANCE=1" "FIRE_B200_B35_BALANCE=0" "FIRE_B200_B35_BALANCE=1"; do
  env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$cfg]', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
done
