#!/bin/bash
OUT=gpurun_out/${1:-r2ab}
mkdir -p $OUT
for a in "fwd 8 256 256 64 32 stats" "fwd 8 256 256 64 32" "dgrad 8 256 256 64 32 mask colsum" "dgrad 8 256 256 64 32" "fwd 8 128 128 128 64 stats" "wgrad 8 256 256 64 32"; do timeout 120 python tools/one_convt.py $a >> $OUT/times.txt 2>&1; done
cat $OUT/times.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_conv_kernel --launch-skip 5 -c 1 -f -o $OUT/convt_fwd python tools/one_convt.py fwd 8 256 256 64 32 stats > $OUT/ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_conv_kernel --launch-skip 5 -c 1 -f -o $OUT/convt_dgrad python tools/one_convt.py dgrad 8 256 256 64 32 mask colsum > $OUT/ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la $OUT
