#!/bin/bash
OUT=gpurun_out/${1:-r2p}
mkdir -p $OUT
for cfg in "fwd 8 512 512 32 32" "fwd 8 512 512 64 32 bits" "dgrad 8 512 512 32 32 bits colsum" "dgrad 8 256 256 64 32 colsum" "fwd 8 512 512 128 32"; do
  timeout 120 python tools/one_op.py $cfg notimeline >> $OUT/one_op.txt 2>&1
  timeout 120 python tools/one_op.py $cfg rowstrip >> $OUT/one_op.txt 2>&1
done
cat $OUT/one_op.txt
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"tc_conv3" --launch-skip 8 -c 2 --csv --log-file $OUT/ncu_rowstrip.csv python tools/one_op.py fwd 8 512 512 32 32 rowstrip > $OUT/ncu.log 2>&1
grep -v "^==" $OUT/ncu_rowstrip.csv | cut -d, -f5,12-15 | tail -12
