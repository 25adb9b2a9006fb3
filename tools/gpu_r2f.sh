#!/bin/bash
# 2-GPU visit: NCCL data-parallel test, bench at N=2 with and without overlapped gradient buckets; plus 1-GPU checks
OUT=gpurun_out/${1:-r2f}
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rP -x > $OUT/pytest_multi.log 2>&1; echo "multi rc=$?" | tee -a $OUT/rc.txt
tail -25 $OUT/pytest_multi.log
timeout 300 python -m pytest tests/test_gpu_unet.py tests/test_gpu_ops.py -m gpu -q -rP -k "side_stream or dense" > $OUT/pytest_sel.log 2>&1; echo "sel rc=$?" | tee -a $OUT/rc.txt
grep -E "passed|failed|side stream|^E  " $OUT/pytest_sel.log | cut -c1-1500 | tail -12
for ov in 1 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --opt comm_overlap=$ov > $OUT/bench_n2_overlap$ov.json 2> $OUT/bench_n2_overlap$ov.err; echo "bench n2 ov=$ov rc=$?" | tee -a $OUT/rc.txt
  head -c 330 $OUT/bench_n2_overlap$ov.json; echo
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$?" | tee -a $OUT/rc.txt
head -c 330 $OUT/bench_n1.json; echo
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 --no-cpu > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
head -c 330 $OUT/bench_cls.json; echo; tail -3 $OUT/bench_cls.err
