#!/bin/bash
OUT=gpurun_out/${1:-r2q}
mkdir -p $OUT
timeout 120 python tools/one_op.py fwd 8 512 512 32 32 rowstrip > $OUT/tl.txt 2>&1
timeout 120 python tools/one_op.py fwd 8 512 512 128 32 rowstrip >> $OUT/tl.txt 2>&1
cat $OUT/tl.txt
