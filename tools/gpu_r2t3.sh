#!/bin/bash
OUT=gpurun_out/${1:-r2t3}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rP -x > $OUT/pytest_multi.log 2>&1; echo "multi rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|B: |dp_worker ok|AssertionError" $OUT/pytest_multi.log | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 40 --warmup 6 --no-cpu > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?" | tee -a $OUT/rc.txt
head -c 330 $OUT/bench_n2.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 40 --warmup 6 --impl reference > $OUT/bench_n2_ref.json 2> $OUT/bench_n2_ref.err; echo "bench n2 ref rc=$?" | tee -a $OUT/rc.txt
head -c 400 $OUT/bench_n2_ref.json; echo
