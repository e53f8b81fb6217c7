#!/bin/bash
# time library variants built into variants/lib*.so on one GPU:  gpurun -- bash scripts/variant_sweep.sh TAG "A B C"
TAG=${1:-variants}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cp dennou-ccm_b200/libdccm_b200.so $OUT/lib_orig.so
for v in ${2:-A B C}; do
  cp variants/lib$v.so dennou-ccm_b200/libdccm_b200.so
  for rep in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT/bench_$v.json") if l.startswith("{")][-1]
    print("variant $v:", round(d["value"], 2), "ex/s", {k: round(x, 3) for k, x in d["part_ms"].items()}, d["output_hash"], d["parity"]["bitwise"])
except Exception as e:
    print("variant $v failed", e); print(open("$OUT/bench_$v.err").read()[-800:])
PY
  done
done
cp $OUT/lib_orig.so dennou-ccm_b200/libdccm_b200.so; rm -f $OUT/lib_orig.so
