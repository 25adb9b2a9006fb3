#!/bin/bash
# quick GPU visit: parity suite + per-op timing table (+ optional A/B of a library option)
OUT=gpurun_out/${1:-q}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/ab_ops.py ${2:+--opt $2} > $OUT/ab_ops.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
tail -32 $OUT/ab_ops.txt
