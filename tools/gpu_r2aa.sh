#!/bin/bash
OUT=gpurun_out/${1:-r2aa}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "wgrad" > $OUT/pytest_wgrad.log 2>&1; echo "wgrad rc=$?" | tee -a $OUT/rc.txt
tail -15 $OUT/pytest_wgrad.log
timeout 600 python tools/ab_ops.py --opt wgrad_dhm=0,1 --kinds conv3x3_wgrad > $OUT/ab_wgrad_dhm.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
grep -E "^conv3x3_wgrad|^step|^op" $OUT/ab_wgrad_dhm.txt
