#!/bin/bash
# pixel-pair stem (Conv2d_1a / 2a): parity, per-op table, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 > gpurun_out/s14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s14_pytest.log
tail -15 gpurun_out/s14_pytest.log
timeout 300 python tools/profile_ops.py 256 512 > gpurun_out/s14_ops.txt 2>&1; head -12 gpurun_out/s14_ops.txt; tail -2 gpurun_out/s14_ops.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-knn --no-frames > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s14_bench.err; head -c 300 gpurun_out/s14_bench.json; echo
