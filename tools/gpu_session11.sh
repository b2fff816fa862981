#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/s11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s11_pytest.log
tail -15 gpurun_out/s11_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s11_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s11_bench.json'))
print(d['value'], d['e2e']['value'])
print({k:(v['us_per_call'], round(v['frac_of_hbm'],3)) for k,v in d['knn']['1M_small_batch'].items() if k.startswith('Q')})
print(d['frames']['upload_routes'], d['frames'].get('parity'))
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/s11_bench.json'))
print(d['encode_small_batch'])
print(d['knn']['1M_small_batch'].get('Q1_host_call'))
PY
