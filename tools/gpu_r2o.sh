#!/bin/bash
OUT=gpurun_out/${1:-r2o}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "rowstrip" > $OUT/pytest_rowstrip.log 2>&1; rc=$?; echo "rowstrip rc=$rc" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^E  |Error" $OUT/pytest_rowstrip.log | cut -c1-300 | tail -12
if [ $rc -ne 0 ]; then tail -30 $OUT/pytest_rowstrip.log | cut -c1-200; exit 0; fi
timeout 300 python tools/ab_ops.py --opt tc_rowstrip=0,1,2 > $OUT/ab_rowstrip.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
grep -E "^conv3x3_(fwd|dgrad)|^step" $OUT/ab_rowstrip.txt | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10}'
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 300 $OUT/bench_nocpu.json; echo
