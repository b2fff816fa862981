#!/bin/bash
# final single-GPU evidence of round 2 (with pool_conv_fused_kernel): full GPU suite, default bench, reference arm, per-op table, launch list of the bench command, block35/block8 ncu
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s34_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s34_pytest.log
tail -4 gpurun_out/s34_pytest.log
timeout 900 python bench.py > gpurun_out/s34_bench.json 2> gpurun_out/s34_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/s34_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s34_bench_reference.json 2> gpurun_out/s34_bench_reference.err
echo "reference rc=$?"; cat gpurun_out/s34_bench_reference.json | head -c 600; echo
timeout 300 python tools/profile_ops.py 256 512 > gpurun_out/s34_ops.txt 2>&1
timeout 300 python tools/ncu_all.py > gpurun_out/s34_ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -o gpurun_out/r02_full -f python tools/ncu_all.py > gpurun_out/s34_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_full.ncu-rep
timeout 300 python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s34_bench_small.json 2>/dev/null &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s34_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
