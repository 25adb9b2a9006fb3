#!/bin/bash
OUT=gpurun_out/${1:-r2x}
mkdir -p $OUT
timeout 600 python tools/ab_ops.py --opt l2_fetch=128,64,32 > $OUT/ab_l2fetch.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
grep -E "maxpool_bwd|convt_|bn_bwd_apply|^step|totals" $OUT/ab_l2fetch.txt | head -40
timeout 600 python -m pytest tests/test_gpu_runners.py -q -x > $OUT/pytest_runners.log 2>&1; echo "runners rc=$?" | tee -a $OUT/rc.txt
tail -4 $OUT/pytest_runners.log
timeout 300 python bench.py --workload classifier224x3 --steps 40 --warmup 5 --no-cpu > $OUT/bench_cls.json 2> $OUT/bench.err; echo "cls rc=$?" | tee -a $OUT/rc.txt
python - $OUT/bench_cls.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"])
PY
