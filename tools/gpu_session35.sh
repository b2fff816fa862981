#!/bin/bash
# strip kernel: patch ring last, stages aligned to the operand swizzle only (Conv2d_2b: 4 -> 6 stages)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "strip or pair_stem or config2 or golden or config1 or pool_conv" > gpurun_out/s35_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s35_pytest.log
tail -6 gpurun_out/s35_pytest.log
FIRE_B200_TRACE_OP=2 timeout 300 python tools/profile_ops.py 256 512 2>&1 | grep -A8 "trace op\|^  [0-2] Conv"
for i in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
done
