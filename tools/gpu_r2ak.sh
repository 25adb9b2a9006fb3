#!/bin/bash
OUT=gpurun_out/${1:-r2ak}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -4 $OUT/pytest.log
timeout 400 python tools/ab_ops.py --opt bn_async=0,1 --kinds bn_apply,bn_bwd_apply,bn_apply_pool,maxpool_bwd > $OUT/ab_bn_async.txt 2>&1
grep -E "^bn_|^maxpool|^step|^op" $OUT/ab_bn_async.txt
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | head -c 240; echo
timeout 300 python bench.py --workload unetpp512 --steps 30 --warmup 5 --no-cpu 2>/dev/null | head -c 240; echo
