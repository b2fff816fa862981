#!/bin/bash
# final evidence pass: smoke, ncu --set full over every kernel family, launch list of the bench command
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s10_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/s10_smoke.log
timeout 300 python tools/ncu_all.py > gpurun_out/s10_ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -o gpurun_out/r02_full -f python tools/ncu_all.py > gpurun_out/s10_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_full.ncu-rep
timeout 300 python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s10_bench_small.json 2>/dev/null &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s10_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | grep -E "r02_ncu|s10"
