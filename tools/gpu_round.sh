#!/bin/bash
# One GPU-box visit: [smoke,] bench line, ncu launch list, ncu --set full of the tensor-core kernels, [GPU parity suite].
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [steps: smoke,bench,list,full,tests]
# gpurun_out/ is capped at 64 MiB: .ncu-rep files are converted to CSV on the box and deleted.
TAG=${1:-r1}
WHAT=${2:-smoke,bench,list,full,tests}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
has() { [[ ",$WHAT," == *",$1,"* ]]; }
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/rc.txt
fi
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 5 --per-op > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
fi
if has list; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu-list rc=$?" | tee -a $OUT/rc.txt
fi
if has full; then
  timeout 600 ncu --set full --clock-control none -k regex:"tc_conv3_kernel|tc_wgrad|tc_conv_kernel" --launch-skip 150 -c 44 \
    -f -o $OUT/tc_full python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $OUT/ncu_full.log 2>&1; echo "ncu-full rc=$?" | tee -a $OUT/rc.txt
  python tools/ncu_summary.py $OUT/tc_full.ncu-rep $OUT/tc_full_summary.csv >> $OUT/ncu_full.log 2>&1
  rm -f $OUT/tc_full.ncu-rep
fi
tail -3 $OUT/smoke.log 2>/dev/null; tail -5 $OUT/pytest_gpu.log 2>/dev/null; head -c 1200 $OUT/bench.json 2>/dev/null
du -sh gpurun_out
