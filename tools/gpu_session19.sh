#!/bin/bash
# N-tile experiments for the M = 2304 layers (Block8 heads: ops 78 + 4j; Mixed_7a 3x3/2: 73, 74, 76)
mkdir -p gpurun_out
H="78 82 86 90 94 98"
mk() { local bn=$1; shift; local s=""; for o in "$@"; do s="$s$o:$bn,"; done; echo "$s"; }
for cfg in "" "$(mk 64 $H)" "$(mk 96 $H)" "$(mk 128 $H)" "$(mk 192 $H)" "73:64,74:64,76:64" "73:96,74:128,76:128" "73:128,74:128,76:128" "73:192,74:256,76:256"; do
  FIRE_B200_BN="$cfg" timeout 300 python bench.py --steps 200 --warmup 10 --no-knn --no-frames --no-cpu --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BN[$cfg]', round(d['value']), d['ms_per_step'], d['parity']['min_cos_vs_fp32_oracle'])"
done 2>&1 | tee gpurun_out/s19_bn.txt
