#!/bin/bash
# side measurements: Dice parity protocol + the other BASELINE configs (no CPU legs)
OUT=gpurun_out/${1:-extra}
mkdir -p $OUT
for wl in unet256 unetpp512; do
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --workload $wl > $OUT/bench_$wl.json 2>$OUT/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"
done
timeout 500 python tools/dice_parity.py > $OUT/dice_parity.json 2>$OUT/dice.err; echo "dice rc=$?"; cat $OUT/dice_parity.json | cut -c1-1500
