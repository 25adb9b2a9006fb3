#!/bin/bash
OUT=gpurun_out/${1:-r2d}
mkdir -p $OUT
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"tc_conv3_kernel" --launch-skip 6 -c 1 -f -o $OUT/halo_32_32 python tools/one_conv.py 8 512 512 32 32 1 > $OUT/ncu_halo.log 2>&1; echo "ncu halo rc=$?" | tee -a $OUT/rc.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"tc_conv3w_kernel" --launch-skip 6 -c 1 -f -o $OUT/dw_32_32 python tools/one_conv.py 8 512 512 32 32 1 dwmerge > $OUT/ncu_dw.log 2>&1; echo "ncu dw rc=$?" | tee -a $OUT/rc.txt
ls -la $OUT
timeout 900 python -m pytest tests/test_gpu_runners.py tests/test_gpu_unet.py -m gpu -q -rP -k "sequential or class_weight or intermediate or emulator" > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|^FAILED|^ERROR|fp16 engine vs" $OUT/pytest_sel.log | tail
