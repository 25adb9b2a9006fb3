#!/bin/bash
OUT=gpurun_out/${1:-r2af}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > $OUT/pytest_tc.log 2>&1; echo "tc rc=$?" | tee -a $OUT/rc.txt
tail -4 $OUT/pytest_tc.log
timeout 600 python tools/ab_ops.py --opt tc_bgroup=1,3 --kinds conv3x3_fwd,conv3x3_dgrad > $OUT/ab_bgroup.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
grep -E "^conv3x3|^step|^op" $OUT/ab_bgroup.txt
