#!/bin/bash
OUT=gpurun_out/${1:-r2r}
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -20
grep -E "^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -10
grep -E "engine fp16 |config\[0\]|side stream|fp16 engine vs" $OUT/pytest_gpu.log | cut -c1-250
for rs in 2 0; do
timeout 300 python bench.py --workload unetpp512 --steps 10 --warmup 3 --no-cpu --opt tc_rowstrip=$rs > $OUT/bench_upp_rs$rs.json 2>$OUT/bench_upp.err; echo "bench-upp rs=$rs rc=$?" | tee -a $OUT/rc.txt
head -c 260 $OUT/bench_upp_rs$rs.json; echo
done
