#!/bin/bash
# inference BN folding + hoisted weight-only ops: op / whole-net parity, classifier bench before / after
OUT=gpurun_out/${1:-r2v}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x -k "post_activation or conv3x3_fwd" > $OUT/pytest_tc.log 2>&1; echo "tc rc=$?" | tee -a $OUT/rc.txt
timeout 900 python -m pytest tests/test_gpu_runners.py tests/test_gpu_unet.py -q -x -k "inference or fp16_inference or intermediate or sequential or golden or facade" > $OUT/pytest_net.log 2>&1; echo "net rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/pytest_tc.log $OUT/pytest_net.log
timeout 300 python bench.py --workload classifier224x3 --steps 40 --warmup 5 --per-op --no-cpu > $OUT/bench_cls.json 2> $OUT/bench.err; echo "cls rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --workload classifier224x3 --steps 40 --warmup 5 --per-op --no-cpu --plan fuse_bn_infer=0 > $OUT/bench_cls_nofold.json 2>> $OUT/bench.err; echo "cls-nofold rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --workload classifier224x3 --steps 40 --warmup 5 --per-op --no-cpu --plan fuse_bn_infer=0,hoist_prep=0 > $OUT/bench_cls_r1.json 2>> $OUT/bench.err; echo "cls-old rc=$?" | tee -a $OUT/rc.txt
for f in bench_cls bench_cls_nofold bench_cls_r1; do python - $OUT/$f.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], {k:v for k,v in d.get("op_breakdown_ms",{}).items() if k!="_per_op"})
PY
done
