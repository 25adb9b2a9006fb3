#!/bin/bash
OUT=gpurun_out/${1:-r2ah}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x -k "multicast" --timeout 60 > $OUT/pytest_mcast.log 2>&1; echo "mcast rc=$?" | tee -a $OUT/rc.txt
tail -8 $OUT/pytest_mcast.log
if grep -q "passed" $OUT/pytest_mcast.log && ! grep -q "failed" $OUT/pytest_mcast.log; then
  for c in 0 2 4 8; do timeout 100 python tools/one_op.py fwd 8 32 32 512 512 notimeline opt:tc_mcast=$c 2>&1 | tail -1; timeout 100 python tools/one_op.py fwd 8 128 128 256 128 notimeline opt:tc_mcast=$c 2>&1 | tail -1; timeout 100 python tools/one_op.py fwd 8 64 64 256 256 stats notimeline opt:tc_mcast=$c 2>&1 | tail -1; timeout 100 python tools/one_op.py fwd 8 64 64 512 256 notimeline opt:tc_mcast=$c 2>&1 | tail -1; done | tee $OUT/one_op_mcast.txt
fi
