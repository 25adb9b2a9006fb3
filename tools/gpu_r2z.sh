#!/bin/bash
# split concat: full GPU suite, bench before / after, per-op table
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -6 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu > $OUT/bench_split.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --plan split_concat=0 > $OUT/bench_nosplit.json 2>> $OUT/bench.err; echo "bench0 rc=$?" | tee -a $OUT/rc.txt
for f in bench_split bench_nosplit; do python - $OUT/$f.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
timeout 600 python tools/ab_ops.py > $OUT/per_op_split.txt 2>&1
timeout 600 python tools/ab_ops.py --plan split_concat=0 > $OUT/per_op_nosplit.txt 2>&1
tail -25 $OUT/per_op_split.txt
