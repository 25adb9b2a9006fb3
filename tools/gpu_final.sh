#!/bin/bash
# final measurements of the round: bench lines (all BASELINE configs), ncu launch list of the bench command, ncu --set full
# of one step's 34 tcgen05 3x3 launches (roofline.traffic) and of the other tensor-core kernels
OUT=gpurun_out/${1:-final}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/rc.txt
timeout 600 python bench.py --steps 40 --warmup 5 --per-op > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 400 $OUT/bench.json; echo
timeout 300 python bench.py --workload unet256 --steps 30 --warmup 5 --no-cpu > $OUT/bench_unet256.json 2>> $OUT/bench.err; echo "bench256 rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --workload unetpp512 --steps 30 --warmup 5 --no-cpu > $OUT/bench_unetpp512.json 2>> $OUT/bench.err; echo "benchpp rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --workload classifier224x3 --steps 40 --warmup 5 --per-op > $OUT/bench_classifier.json 2>> $OUT/bench.err; echo "benchcls rc=$?" | tee -a $OUT/rc.txt
for f in bench_unet256 bench_unetpp512 bench_classifier; do head -c 200 $OUT/$f.json; echo; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 900 -c 200 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu-list rc=$?" | tee -a $OUT/rc.txt
timeout 600 ncu --set full --clock-control none -k regex:"tc_conv3" --launch-skip 102 -c 34 \
  -f -o $OUT/tc3_full python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $OUT/ncu_full.log 2>&1; echo "ncu-full rc=$?" | tee -a $OUT/rc.txt
python tools/ncu_summary.py $OUT/tc3_full.ncu-rep $OUT/tc3_full_summary.csv >> $OUT/ncu_full.log 2>&1
rm -f $OUT/tc3_full.ncu-rep
du -sh gpurun_out
