#!/bin/bash
# CTA pairs for the layers with 64 tile pairs (Mixed_6a b1 3x3/2, Mixed_7a b2 3x3: one 128 x 256 tile per CTA, 0.9-1.2 MB of weights each)
mkdir -p gpurun_out
for cfg in "FIRE_B200_PAIR_SLACK=0" "FIRE_B200_PAIR_SLACK=10" "FIRE_B200_PAIR_SLACK=0" "FIRE_B200_PAIR_SLACK=10"; do
  env $cfg timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$cfg]', d['value'], d['ms_per_step'], d['parity'], {k: round(v,4) for k,v in d['roofline']['families']['by_stage_ms_serialised'].items() if k in ('Mixed_6a','Mixed_7a')})"
done 2>&1 | tee gpurun_out/s32_pair_slack.txt
FIRE_B200_PAIR_SLACK=10 timeout 300 python tools/profile_ops.py 256 512 2>&1 | grep -n "Mixed" 
