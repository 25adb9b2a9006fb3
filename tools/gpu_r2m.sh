#!/bin/bash
OUT=gpurun_out/${1:-r2m}
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -rP --deselect tests/test_gpu_dice.py::test_dice_protocol_means_vs_oracle_fixture > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -20
grep -E "^E  " $OUT/pytest_gpu.log | cut -c1-400 | head -10
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 --per-op > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
python - <<PY
import json
d=json.load(open("$OUT/bench_cls.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["auroc_vs_oracle"], d["cpu_baseline"])
for r in d["op_breakdown_ms"]["_per_op"][:3]: print(r)
PY
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 300 $OUT/bench_nocpu.json; echo
