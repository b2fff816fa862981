#!/bin/bash
# strip kernel phase accounting for the three stem convs
mkdir -p gpurun_out
for op in 0 1 2; do
  FIRE_B200_TRACE_OP=$op timeout 300 python tools/profile_ops.py 256 512 2> gpurun_out/s15_trace_op$op.txt > /dev/null
  echo "== op $op"; tail -22 gpurun_out/s15_trace_op$op.txt
done
