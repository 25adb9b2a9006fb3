#!/bin/bash
OUT=gpurun_out/${1:-r2l}
mkdir -p $OUT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"maxpool_bwd|head_bwd|c1_fwd|c1_wgrad" --launch-skip 14 -c 7 -f -o $OUT/tail python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > $OUT/ncu_tail.log 2>&1; echo "ncu tail rc=$?" | tee -a $OUT/rc.txt
ls -la $OUT
timeout 300 python -m pytest tests/test_gpu_preprocess.py -m gpu -q > $OUT/pytest_pp.log 2>&1; echo "pp rc=$?" | tee -a $OUT/rc.txt
tail -3 $OUT/pytest_pp.log
