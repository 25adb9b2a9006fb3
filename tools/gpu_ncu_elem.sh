#!/bin/bash
# ncu --set full of the HBM-bound (non-tensor-core) kernels of one training step -> CSV summary (rep deleted: 64 MiB cap)
OUT=gpurun_out/${1:-elem}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none -k regex:"maxpool|bn_reduce|bn_apply|bn_bwd|conv3x3_c1|head_|adam_kernel|channel_sum|pack" \
  --launch-skip 200 -c 70 -f -o $OUT/elem_full python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $OUT/ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py $OUT/elem_full.ncu-rep $OUT/elem_full_summary.csv
rm -f $OUT/elem_full.ncu-rep
du -sh gpurun_out
