#!/bin/bash
# round-2 first GPU visit: validate the never-run kernels (1-bit ReLU masks, dw-merged thin layers), full suite, A/B
OUT=gpurun_out/${1:-r2a}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "relu_bits or dwmerge" > $OUT/new_kernels.log 2>&1; echo "new-kernels rc=$?" | tee -a $OUT/rc.txt
tail -30 $OUT/new_kernels.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tc.py::test_tc_conv3x3_relu_bits_roundtrip --deselect tests/test_gpu_tc.py::test_tc_conv3x3_dwmerge_fwd_and_dgrad > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/ab_ops.py --opt tc_dwmerge=0,1 > $OUT/ab_dwmerge.txt 2>&1; echo "ab-dwmerge rc=$?" | tee -a $OUT/rc.txt
tail -26 $OUT/ab_dwmerge.txt
timeout 300 python tools/ab_ops.py --plan relu_bits=1 > $OUT/ab_relubits.txt 2>&1; echo "ab-relubits rc=$?" | tee -a $OUT/rc.txt
tail -26 $OUT/ab_relubits.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_nocpu.json 2>$OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
head -c 600 $OUT/bench_nocpu.json; echo
timeout 300 python -m pytest tests/test_gpu_unet.py -m gpu -q -s -k "side_stream" > $OUT/side.log 2>&1; echo "side rc=$?" | tee -a $OUT/rc.txt
grep -i "side stream\|passed\|failed" $OUT/side.log | tail
