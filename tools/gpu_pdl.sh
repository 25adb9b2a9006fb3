#!/bin/bash
OUT=gpurun_out/${1:-pdl}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -4 $OUT/pytest_gpu.log
for v in 0 1; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --opt pdl=$v > $OUT/bench_pdl$v.json 2>$OUT/bench$v.err; echo "bench pdl=$v rc=$?"
  python -c "import json;d=json.load(open('$OUT/bench_pdl$v.json'));print('pdl=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['last_loss_dice'])"
done
