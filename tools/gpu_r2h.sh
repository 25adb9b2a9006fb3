#!/bin/bash
# 2-GPU visit: NCCL data-parallel pytest; new 1-GPU tests (runners / HDF5 / preprocess)
OUT=gpurun_out/${1:-r2h}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rP -x > $OUT/pytest_multi.log 2>&1; echo "multi rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|B: sync|dp_worker ok|rank.*Error|AssertionError" $OUT/pytest_multi.log | tail -12
timeout 600 python -m pytest tests/test_gpu_runners.py tests/test_gpu_preprocess.py -m gpu -q > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^E  " $OUT/pytest_sel.log | cut -c1-400 | tail -12
