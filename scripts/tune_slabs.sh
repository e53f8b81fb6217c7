#!/bin/bash
# slab-pipeline sweep on one GPU:  gpurun --timeout 900 -- bash scripts/tune_slabs.sh TAG "0 8 16 32 64"
TAG=${1:-slabs}; OUT=gpurun_out/$TAG; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_zz_late_gpu.py -m gpu -q -x ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for n in ${2:-0 8 16 32 64}; do
  ( timeout 300 python bench.py --steps 20 --warmup 3 --slabs $n --no-e2e --no-cpu --no-others ) > $OUT/bench_s$n.json 2> $OUT/bench_s$n.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT/bench_s$n.json") if l.startswith("{")][-1]
    print("slabs $n:", round(d["value"], 2), "exchanges/s", round(d["ms_per_step"], 3), "ms", d["output_hash"], "parity bitwise", d.get("parity", {}).get("bitwise"), d.get("parity", {}).get("error"))
except Exception as e:
    print("slabs $n failed", e); print(open("$OUT/bench_s$n.err").read()[-1500:])
PY
done
