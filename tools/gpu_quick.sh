#!/bin/bash
# quick GPU visit: [a canary test first,] parity suite + per-op timing table (+ optional A/B of a library option)
# usage: bash tools/gpu_quick.sh <tag> [opt=v1,v2] [canary -k expression]
OUT=gpurun_out/${1:-q}
mkdir -p $OUT
if [ -n "$3" ]; then
  timeout 180 python -m pytest tests -m gpu -x -q -k "$3" > $OUT/canary.log 2>&1; rc=$?
  echo "canary rc=$rc" | tee -a $OUT/rc.txt; tail -5 $OUT/canary.log
  if [ $rc -ne 0 ]; then tail -40 $OUT/canary.log; exit 1; fi
fi
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/ab_ops.py ${2:+--opt $2} > $OUT/ab_ops.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
tail -24 $OUT/ab_ops.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 400 $OUT/bench_nocpu.json; echo
