#!/bin/bash
# GPU session: full GPU test suite, then the bench at the driver's settings.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/s1_bench.json
