#!/bin/bash
OUT=gpurun_out/${1:-r2c}
mkdir -p $OUT
for cfg in "8 512 512 32 32" "8 256 256 64 64" "8 512 512 64 32"; do
  timeout 120 python tools/one_conv.py $cfg 1 timeline >> $OUT/timeline_halo.txt 2>&1
  timeout 120 python tools/one_conv.py $cfg 1 timeline dwmerge >> $OUT/timeline_dwmerge.txt 2>&1
done
cat $OUT/timeline_halo.txt | head -40
cat $OUT/timeline_dwmerge.txt | head -40
timeout 900 python -m pytest tests/test_gpu_runners.py -m gpu -q -x > $OUT/pytest_runners.log 2>&1; echo "runners rc=$?" | tee -a $OUT/rc.txt
tail -30 $OUT/pytest_runners.log
