#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 200 --warmup 10 > gpurun_out/s38_bench_n4.json 2> gpurun_out/s38_bench_n4.err
echo "bench n4 rc=$?"; tail -2 gpurun_out/s38_bench_n4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s38_bench_n4.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'])
k=d['knn']['10M']; print(k['value'], k['ms_per_batch'], k['roofline']['frac'], k.get('parity',{}).get('ids_equal_except_1e-5_ties'))
print(d['frames']['faces_per_s'])
PY
