#!/bin/bash
# Multi-GPU visit: gpurun --gpus N -- bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tests/sharded_gpu_check.py T106_1deg ) > $OUT/check_T106.log 2>&1; echo "check T106 exit $?" | tee -a $OUT/check_T106.log
( timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    tests/sharded_gpu_check.py T341_0p25deg ) > $OUT/check_T341.log 2>&1; echo "check T341 exit $?" | tee -a $OUT/check_T341.log
for h in nccl allgather; do
( DCCM_HALO=$h timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    tests/sharded_gpu_check.py T106_1deg ) > $OUT/check_T106_$h.log 2>&1; echo "check T106 halo=$h exit $?" | tee -a $OUT/check_T106_$h.log
done
grep "rank" $OUT/check_T106*.log $OUT/check_T341.log | head -40
n=${3:-1}
while [ $n -le $N ]; do
  if [ $n -eq 1 ]; then
    ( timeout -k 5 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    ( timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620+n)) \
        bench.py --gpus $n --steps 10 --warmup 3 ) > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  fi
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$OUT/bench_n$n.json') if l.startswith('{')][-1]
    print('N=$n', round(d['value'],2), 'ex/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['part_ms'].items()}, 'e2e', d.get('e2e',{}).get('value'))
except Exception as e:
    print('N=$n failed', e); print(open('$OUT/bench_n$n.err').read()[-1500:])
PY
  n=$((n*2))
done
# BASELINE config 4: T341 <-> 0.25 deg sharded over 2 / 4 GPUs
for n in 2 4; do
  [ $n -le $N ] || continue
  ( timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29640+n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --workload T341_0p25deg --no-e2e ) > $OUT/bench_T341_n$n.json 2> $OUT/bench_T341_n$n.err
  python -c "
import json
try:
    d=[json.loads(l) for l in open('$OUT/bench_T341_n$n.json') if l.startswith('{')][-1]
    print('T341 N=$n', round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms')
except Exception as e:
    print('T341 N=$n failed', e)
"
done
# BASELINE config 2 across GPUs: the 64-member ensemble sharded by member (replicas of the tables, no collective)
( timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29660 \
    bench.py --gpus $N --steps 20 --warmup 3 --workload T42x64 --no-e2e ) > $OUT/bench_T42x64_n$N.json 2> $OUT/bench_T42x64_n$N.err
cut -c1-400 $OUT/bench_T42x64_n$N.json
# the other two halo transports at N (default above is the peer-memory halo)
for h in nccl allgather; do
  ( timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29670 \
      bench.py --gpus $N --steps 10 --warmup 3 --halo $h --no-e2e ) > $OUT/bench_halo_${h}_n$N.json 2> $OUT/bench_halo_${h}_n$N.err
  python -c "
import json
try:
    d=[json.loads(l) for l in open('$OUT/bench_halo_${h}_n$N.json') if l.startswith('{')][-1]
    print('halo $h N=$N', round(d['value'],1), 'ex/s')
except Exception as e:
    print('halo $h N=$N failed', e)
"
done
