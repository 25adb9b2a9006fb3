#!/bin/bash
OUT=gpurun_out/${1:-r2u}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_unet.py -m gpu -q -k "switched_off or gamma" > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^E    " $OUT/pytest_sel.log | cut -c1-260 | tail -24
