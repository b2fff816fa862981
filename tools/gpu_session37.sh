#!/bin/bash
# closing record of round 2: full GPU suite, default bench, per-op table, launch list of the bench command
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s37_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s37_pytest.log
tail -3 gpurun_out/s37_pytest.log
timeout 900 python bench.py > gpurun_out/s37_bench.json 2> gpurun_out/s37_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/s37_bench.json; echo
timeout 300 python tools/profile_ops.py 256 512 > gpurun_out/s37_ops.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s37_bench_small.json 2>/dev/null &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s37_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
