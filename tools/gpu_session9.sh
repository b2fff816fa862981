#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
tail -6 gpurun_out/s9_pytest.log
for tw in 1 0; do for pdl in 1; do
  FIRE_B200_TAIL_WAIT=$tw timeout 300 python bench.py --steps 400 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('TAIL_WAIT=$tw', d['value'], d['ms_per_step'], d['parity'])" >> gpurun_out/s9_tail.txt
done; done
FIRE_B200_TAIL_WAIT=0 timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=600 > gpurun_out/s9_pytest_tail0.log 2>&1
echo "pytest(tail0) rc=$?" >> gpurun_out/s9_pytest_tail0.log; tail -4 gpurun_out/s9_pytest_tail0.log
cat gpurun_out/s9_tail.txt
