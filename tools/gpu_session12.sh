#!/bin/bash
# first contact of block8_fused_kernel: parity vs the per-layer path, timeline, per-op table, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "block8 or config2 or saturates" > gpurun_out/s12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s12_pytest.log
tail -25 gpurun_out/s12_pytest.log
timeout 300 python tools/trace_block8.py 256 2> gpurun_out/s12_block8_timeline.txt; tail -8 gpurun_out/s12_block8_timeline.txt
timeout 300 python tools/profile_ops.py 256 512 > gpurun_out/s12_ops.txt 2>&1; tail -32 gpurun_out/s12_ops.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-knn --no-frames > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s12_bench.err; head -c 600 gpurun_out/s12_bench.json
