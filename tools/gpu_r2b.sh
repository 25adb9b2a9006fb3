#!/bin/bash
# round-2 GPU visit: full parity suite (no -x, prints of passed tests kept), dw-merge A/B per op, bench line
OUT=gpurun_out/${1:-r2b}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -30
grep -E "fp16 engine vs|unet 512 b8|side stream:" $OUT/pytest_gpu.log
timeout 300 python tools/ab_ops.py --opt tc_dwmerge=0,1 > $OUT/ab_dwmerge.txt 2>&1; echo "ab-dwmerge rc=$?" | tee -a $OUT/rc.txt
grep -E "^conv3x3_(fwd|dgrad)|^step" $OUT/ab_dwmerge.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 400 $OUT/bench_nocpu.json; echo
