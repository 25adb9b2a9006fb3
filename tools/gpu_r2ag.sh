#!/bin/bash
OUT=gpurun_out/${1:-r2ag}
mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_tc.py -q -x -k "multicast" --timeout 60 > $OUT/pytest_mcast.log 2>&1; echo "mcast rc=$?" | tee -a $OUT/rc.txt
tail -15 $OUT/pytest_mcast.log
if grep -q "passed" $OUT/pytest_mcast.log && ! grep -q "failed" $OUT/pytest_mcast.log; then
  for c in 0 1; do timeout 100 python tools/one_op.py fwd 32 32 32 512 512 notimeline opt:tc_mcast=$c 2>&1 | tail -1; timeout 100 python tools/one_op.py fwd 8 128 128 256 128 notimeline opt:tc_mcast=$c 2>&1 | tail -1; timeout 100 python tools/one_op.py fwd 8 64 64 256 256 stats notimeline opt:tc_mcast=$c 2>&1 | tail -1; done
  timeout 400 python tools/ab_ops.py --opt tc_mcast=0,1 --kinds conv3x3_fwd,conv3x3_dgrad > $OUT/ab_mcast.txt 2>&1; echo "ab rc=$?" | tee -a $OUT/rc.txt
  grep -E "^conv3x3|^step|^op" $OUT/ab_mcast.txt
fi
