#!/bin/bash
# GPU session 2: tests, bench, a tile experiment, ncu --set full over every kernel family, launch list of the bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -4 gpurun_out/s2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
echo "bench rc=$?"
for cfg in "" "81:256,85:256,89:256,93:256,97:256,101:256"; do
  FIRE_B200_BN="$cfg" timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BN[$cfg]', d['value'], d['ms_per_step'], d['parity'])" >> gpurun_out/s2_bn_experiment.txt
done
cat gpurun_out/s2_bn_experiment.txt
FIRE_B200_BN="81:256,85:256,89:256,93:256,97:256,101:256" timeout 300 python tools/profile_ops.py > gpurun_out/s2_ops_up256.txt 2>&1
timeout 300 python tools/profile_ops.py > gpurun_out/s2_ops_default.txt 2>&1
timeout 300 python tools/ncu_all.py > gpurun_out/s2_ncu_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -o gpurun_out/r02_full -f python tools/ncu_all.py > gpurun_out/s2_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r02_full.ncu-rep
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_full.ncu-rep
timeout 300 python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s2_bench_small.json 2>/dev/null &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-knn --no-frames --no-cpu --no-sustained > gpurun_out/s2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | tail -20
