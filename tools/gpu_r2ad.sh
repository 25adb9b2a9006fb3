#!/bin/bash
OUT=gpurun_out/${1:-r2ad}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "batchnorm or dropout" > $OUT/pytest_ops.log 2>&1; echo "ops rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_runners.py -q -x -k "unetpp or golden or runner" > $OUT/pytest_upp.log 2>&1; echo "upp rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/pytest_upp.log
timeout 300 python bench.py --workload unetpp512 --steps 30 --warmup 5 --no-cpu > $OUT/bench_upp.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --workload unetpp512 --steps 30 --warmup 5 --no-cpu --plan fuse_dropout_bn=0 > $OUT/bench_upp_off.json 2>> $OUT/bench.err; echo "bench0 rc=$?" | tee -a $OUT/rc.txt
for f in bench_upp bench_upp_off; do python - $OUT/$f.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["config"].get("last_loss_dice"))
PY
done
timeout 400 python tools/ab_ops.py --workload unetpp512 > $OUT/per_op.txt 2>&1; tail -28 $OUT/per_op.txt
