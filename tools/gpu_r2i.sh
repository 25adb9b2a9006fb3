#!/bin/bash
OUT=gpurun_out/${1:-r2i}
mkdir -p $OUT
for cfg in "fwd 8 512 512 32 32 stats" "fwd 8 512 512 32 32 stats dwmerge" "fwd 8 512 512 64 32 bits" "fwd 8 512 512 64 32 bits dwmerge" \
           "dgrad 8 512 512 32 32 bits colsum" "dgrad 8 512 512 32 32 bits colsum dwmerge" "dgrad 8 512 512 32 64 colsum" "dgrad 8 512 512 32 64 colsum dwmerge" \
           "dgrad 8 256 256 64 128 colsum" "fwd 8 256 256 128 64 bits" "fwd 8 256 256 64 64 stats" "dgrad 8 256 256 64 64 bits colsum"; do
  timeout 120 python tools/one_op.py $cfg >> $OUT/one_op.txt 2>&1
done
cat $OUT/one_op.txt
timeout 300 python bench.py --workload classifier224x3 --steps 20 --warmup 5 --no-cpu --per-op > $OUT/bench_cls.json 2>$OUT/bench_cls.err; echo "bench-cls rc=$?" | tee -a $OUT/rc.txt
python - <<PY
import json
d=json.load(open("$OUT/bench_cls.json"))
print(d["value"], d["ms_per_step"])
for r in d["op_breakdown_ms"]["_per_op"]: print(r)
PY
