#!/bin/bash
OUT=gpurun_out/${1:-r2n}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dice.py tests/test_gpu_ops.py tests/test_gpu_runners.py -m gpu -q -rP > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|engine fp16|config\[0\]" $OUT/pytest_sel.log | cut -c1-300 | tail -12
timeout 300 python tools/ab_ops.py > $OUT/per_op.txt 2>&1; echo "per-op rc=$?" | tee -a $OUT/rc.txt
grep -E "conv2d_1 |head_bwd|^step|maxpool_bwd" $OUT/per_op.txt
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 --no-cpu --per-op > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
python - <<PY
import json
d=json.load(open("$OUT/bench_cls.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 300 $OUT/bench_nocpu.json; echo
