#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s17_pytest.log
tail -5 gpurun_out/s17_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s17_bench.err; head -c 300 gpurun_out/s17_bench.json; echo
timeout 300 python tools/profile_ops.py 256 512 > gpurun_out/s17_ops.txt 2>&1
