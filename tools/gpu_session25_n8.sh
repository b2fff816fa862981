#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/s25_bench_n8.json 2> gpurun_out/s25_bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/s25_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s25_bench_n8.json'))
print(d['value'], d['e2e']['value'])
k=d['knn']['10M']; print(k['value'], k['ms_per_batch'], k['roofline']['frac'], k.get('parity'))
print(d['frames']['faces_per_s'])
PY
