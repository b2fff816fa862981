#!/bin/bash
# block8_fused_kernel with the folded average pool: parity, timeline, bench, ncu --set full of one launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 > gpurun_out/s13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
tail -5 gpurun_out/s13_pytest.log
timeout 300 python tools/trace_block8.py 256 2> gpurun_out/s13_block8_timeline.txt; tail -8 gpurun_out/s13_block8_timeline.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-knn --no-frames > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s13_bench.err; head -c 400 gpurun_out/s13_bench.json; echo
timeout 300 python tools/ncu_forward.py > gpurun_out/s13_fwd_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block8_fused -s 6 -c 1 -o gpurun_out/r02_block8_fused -f python tools/ncu_forward.py > gpurun_out/s13_ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_block8_fused.ncu-rep
