#!/bin/bash
OUT=gpurun_out/${1:-r2e}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -30
timeout 300 python tools/ab_ops.py --opt tc_dwmerge=0,1,2 > $OUT/ab_dwmerge_epi8.txt 2>&1; echo "ab1 rc=$?" | tee -a $OUT/rc.txt
timeout 300 python tools/ab_ops.py --opt tc_dwmerge=0,1,2 --set tc_dw_epi8=0 > $OUT/ab_dwmerge_epi4.txt 2>&1; echo "ab2 rc=$?" | tee -a $OUT/rc.txt
paste -d' ' <(grep -E "^conv3x3_(fwd|dgrad)|^step" $OUT/ab_dwmerge_epi8.txt | awk '{print $1,$2,$3,$4,$5}') <(grep -E "^conv3x3_(fwd|dgrad)|^step" $OUT/ab_dwmerge_epi4.txt | awk '{print $4}')
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 300 $OUT/bench_nocpu.json; echo
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
cat $OUT/bench_cls.json | head -c 3000; tail -3 $OUT/bench_cls.err
timeout 300 python bench.py --workload unetpp512 --steps 10 --warmup 3 --no-cpu --per-op > $OUT/bench_upp.json 2>$OUT/bench_upp.err; echo "bench-upp rc=$?" | tee -a $OUT/rc.txt
head -c 300 $OUT/bench_upp.json; echo
timeout 300 python bench.py --workload unetpp512 --steps 10 --warmup 3 --no-cpu --opt tc_dwmerge=0 > $OUT/bench_upp_dw0.json 2>>$OUT/bench_upp.err; echo "bench-upp0 rc=$?" | tee -a $OUT/rc.txt
head -c 300 $OUT/bench_upp_dw0.json; echo
