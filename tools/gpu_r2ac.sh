#!/bin/bash
OUT=gpurun_out/${1:-r2ac}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "convt or colsum or streamed or stats" > $OUT/pytest_convt.log 2>&1; echo "convt rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/pytest_convt.log
for a in "fwd 8 256 256 64 32 stats" "dgrad 8 256 256 64 32 mask colsum" "fwd 8 128 128 128 64 stats" "dgrad 8 128 128 128 64 mask colsum"; do for jt in 128 64; do timeout 120 python tools/one_convt.py $a opt:convt_jt=$jt >> $OUT/times.txt 2>&1; done; done
cat $OUT/times.txt
timeout 600 python tools/ab_ops.py --opt convt_jt=128,64 --kinds convt_fwd,convt_dgrad > $OUT/ab_convt_jt.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
grep -E "^convt|^step|^op" $OUT/ab_convt_jt.txt
