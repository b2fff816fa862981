#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -x --timeout=600 -k two_gpu > gpurun_out/s24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s24_pytest.log
tail -4 gpurun_out/s24_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/s24_bench_n2.json 2> gpurun_out/s24_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/s24_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s24_bench_n2.json'))
print(d['value'], d['e2e']['value'])
k=d['knn']['10M']; print(k['value'], k['ms_per_batch'], k['roofline']['frac'], k.get('parity'))
print(d['frames']['faces_per_s'])
PY
