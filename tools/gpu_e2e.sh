#!/bin/bash
OUT=gpurun_out/${1:-e2e}
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "pipelined" > $OUT/canary.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -5 $OUT/canary.log
if [ $rc -ne 0 ]; then tail -40 $OUT/canary.log; exit 1; fi
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu > $OUT/bench.json 2>$OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python -c "import json;d=json.load(open('$OUT/bench.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
