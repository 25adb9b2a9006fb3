#!/bin/bash
# A/B of one library option over the per-op timing table (no tests): bash tools/gpu_ab.sh <tag> <opt=v1,v2> [kinds]
OUT=gpurun_out/${1:-ab}
mkdir -p $OUT
timeout 300 python tools/ab_ops.py --opt $2 ${3:+--kinds $3} > $OUT/ab_ops.txt 2>&1; echo "ab rc=$?"
cat $OUT/ab_ops.txt | tail -70
