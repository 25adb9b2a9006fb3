#!/bin/bash
OUT=gpurun_out/${1:-r2j}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_preprocess.py -m gpu -q > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^E  " $OUT/pytest_sel.log | cut -c1-300 | tail -12
timeout 300 python tools/ab_ops.py --opt tc_dwmerge=2,3 > $OUT/ab_dwmerge23.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
grep -E "^conv3x3_(fwd|dgrad)|^step" $OUT/ab_dwmerge23.txt | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9}'
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 --no-cpu --per-op > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
python - <<PY
import json
d=json.load(open("$OUT/bench_cls.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for r in d["op_breakdown_ms"]["_per_op"][:4]: print(r)
PY
for v in 2 3; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --opt tc_dwmerge=$v > $OUT/bench_dw$v.json 2>$OUT/bench.err; echo "bench dw$v rc=$?" | tee -a $OUT/rc.txt
head -c 250 $OUT/bench_dw$v.json; echo
done
