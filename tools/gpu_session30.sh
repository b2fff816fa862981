#!/bin/bash
# last check of the round's final commit: smoke(), the full GPU suite, a short bench
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s30_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/s30_smoke.log
timeout 1700 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s30_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s30_pytest.log
tail -4 gpurun_out/s30_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-knn --no-frames > gpurun_out/s30_bench.json 2> gpurun_out/s30_bench.err
echo "bench rc=$?"; head -c 260 gpurun_out/s30_bench.json; echo
