#!/bin/bash
OUT=gpurun_out/${1:-side}
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "side_stream" > $OUT/canary.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -5 $OUT/canary.log
if [ $rc -ne 0 ]; then tail -40 $OUT/canary.log; exit 1; fi
for v in 0 1; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --opt side_stream=$v > $OUT/bench_side$v.json 2>$OUT/bench$v.err; echo "bench side=$v rc=$?"
  python -c "import json;d=json.load(open('$OUT/bench_side$v.json'));print('side=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['last_loss_dice'])"
done
