#!/bin/bash
OUT=gpurun_out/${1:-r2g}
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -rP --deselect tests/test_gpu_dice.py::test_dice_protocol_means_vs_oracle_fixture > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -30
grep -E "fp16 engine vs|unet 512 b8|side stream:|config\[0\]" $OUT/pytest_gpu.log
grep -E "^E  " $OUT/pytest_gpu.log | cut -c1-600 | head -20
