#!/bin/bash
# N-GPU visit: gradient-exchange A/B (bucketed + overlapped, bucketed serial, one all-reduce) at N = $2 GPUs
OUT=gpurun_out/${1:-scale}
N=${2:-8}
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
run() {  # tag, extra args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 40 --warmup 6 --no-cpu $2 > $OUT/bench_n${N}_$1.json 2> $OUT/bench_n${N}_$1.err; echo "bench n$N $1 rc=$?" | tee -a $OUT/rc.txt
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_n${N}_$1.json"))
    print("$1", "value %.1f ms/step %.4f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$1 failed", e)
PY
}
run buckets_overlap ""
run single "--plan grad_bucket_bytes=0"
run buckets_serial "--opt comm_overlap=0"
run buckets16_overlap "--plan grad_bucket_bytes=16777216"
run buckets_overlap_2 ""
run single_2 "--plan grad_bucket_bytes=0"
