#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "strip or pair_stem or config2 or golden" > gpurun_out/s36_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s36_pytest.log
tail -3 gpurun_out/s36_pytest.log
for op in 0 2; do FIRE_B200_TRACE_OP=$op timeout 300 python tools/profile_ops.py 256 512 2>&1 | grep -A3 "trace op" | head -3; done
timeout 300 python tools/profile_ops.py 256 512 2>&1 | sed -n 3,5p
for i in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
done
