#!/bin/bash
# neighbour handshakes vs all-rank barrier in the peer-memory halo:  gpurun --gpus N -- bash scripts/gpu_sync_ab.sh TAG N
TAG=${1:-sync}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for s in neighbour barrier; do
  ( DCCM_HALO=peer DCCM_SYNC=$s timeout -k 5 240 $TR --master-port 29711 tests/sharded_gpu_check.py T106_1deg ) > $OUT/check_$s.log 2>&1
  echo "check sync=$s exit $?"; grep -c "mismatches \[\]" $OUT/check_$s.log
done
for s in neighbour barrier neighbour barrier; do
  ( timeout -k 5 300 $TR --master-port 29712 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --peer-sync $s ) > $OUT/bench_$s.json 2> $OUT/bench_$s.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT/bench_$s.json") if l.startswith("{")][-1]
    print("sync=$s N=$N:", round(d["value"], 1), "ex/s", round(d["ms_per_step"], 4), "ms", {k: round(v, 3) for k, v in d["part_ms"].items()}, d["output_hash"], d["parity"]["bitwise"])
except Exception as e:
    print("sync=$s failed", e); print(open("$OUT/bench_$s.err").read()[-1200:])
PY
done
