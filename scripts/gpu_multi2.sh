#!/bin/bash
# Round-2 multi-GPU visit:  gpurun --gpus N --timeout T -- bash scripts/gpu_multi2.sh TAG N
# N=2: sharded bit-identity checks (3 halo transports) + T1279 / T341 bench;  N=4: T1279 / T341 bench;
# N=8: bit-identity check, T1279 bench with the three halo transports, 64-member ensemble by member.
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/host.txt; numactl -H >> $OUT/host.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python - "$1" "$2" <<'PY'
import json, sys
f, tag = sys.argv[1], sys.argv[2]
try:
    d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
    p = d.get("parity", {})
    print(tag, round(d["value"], 2), "ex/s", round(d["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in d["part_ms"].items()},
          "hash", d.get("output_hash"), "parity", p.get("bitwise"), p.get("atm_rows"), p.get("error"),
          "setup_s", d["config"].get("setup_s"), "e2e", round(d.get("e2e", {}).get("value", 0), 2), d.get("e2e", {}).get("numa_node"),
          "roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(tag, "failed", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
}
check() {   # workload halo port
  ( DCCM_HALO=$2 timeout -k 5 240 $TR --master-port $3 tests/sharded_gpu_check.py $1 ) > $OUT/check_$1_$2.log 2>&1
  echo "check $1 halo=$2 exit $?" | tee -a $OUT/check_$1_$2.log; grep "mismatches" $OUT/check_$1_$2.log | head -8
}
if [ $N -eq 2 ]; then
  check T106_1deg peer 29611; check T106_1deg nccl 29612; check T106_1deg allgather 29613; check T341_0p25deg peer 29614
fi
if [ $N -eq 8 ]; then check T341_0p25deg peer 29615; fi
( timeout -k 5 400 $TR --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
show $OUT/bench_n$N.json "T1279 N=$N peer"
if [ $N -le 4 ]; then
  ( timeout -k 5 300 $TR --master-port 29622 bench.py --gpus $N --steps 20 --warmup 3 --workload T341_0p25deg ) > $OUT/bench_T341_n$N.json 2> $OUT/bench_T341_n$N.err
  show $OUT/bench_T341_n$N.json "T341 N=$N peer"
fi
if [ $N -eq 8 ]; then
  for h in nccl allgather; do
    ( timeout -k 5 300 $TR --master-port 29623 bench.py --gpus $N --steps 10 --warmup 3 --halo $h --no-e2e ) > $OUT/bench_halo_${h}_n$N.json 2> $OUT/bench_halo_${h}_n$N.err
    show $OUT/bench_halo_${h}_n$N.json "T1279 N=$N halo=$h"
  done
  ( timeout -k 5 300 $TR --master-port 29624 bench.py --gpus $N --steps 20 --warmup 3 --workload T42x64 --no-e2e ) > $OUT/bench_T42x64_n$N.json 2> $OUT/bench_T42x64_n$N.err
  show $OUT/bench_T42x64_n$N.json "T42x64 N=$N by member"
  ( timeout -k 5 200 $TR --master-port 29625 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-full-grid ) > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
  python -c "
import json
d=[json.loads(l) for l in open('$OUT/bench_ref_n$N.json') if l.startswith('{')][-1]; print('reference arm under torchrun N=$N:', d['value'], 'cores', d['cpu_baseline']['cores'])"
fi
