#!/bin/bash
OUT=gpurun_out/${1:-r2k}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_preprocess.py -m gpu -q > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED" $OUT/pytest_sel.log | cut -c1-300 | tail -12
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 --no-cpu --per-op > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
python - <<PY
import json
d=json.load(open("$OUT/bench_cls.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for r in d["op_breakdown_ms"]["_per_op"][:3]: print(r)
PY
