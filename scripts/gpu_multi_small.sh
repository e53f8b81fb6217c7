#!/bin/bash
# overhead study on 2 GPUs with the T341 workload (per-rank work ~ T1279 on 16 ranks)
OUT=gpurun_out/${1:-ms}; mkdir -p $OUT
for mode in "" "--no-graph"; do
for n in 1 2; do
  tag=n${n}$(echo $mode | tr -d ' -')
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --workload T341_0p25deg --steps 20 --warmup 3 --no-cpu --no-e2e $mode > $OUT/b_$tag.json 2> $OUT/b_$tag.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) \
       bench.py --gpus $n --workload T341_0p25deg --steps 20 --warmup 3 --no-e2e $mode > $OUT/b_$tag.json 2> $OUT/b_$tag.err
  fi
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$OUT/b_$tag.json') if l.startswith('{')][-1]
    print('$tag', round(d['value'],1), 'ex/s', round(d['ms_per_step'],4), 'ms', {k: round(v,3) for k,v in d['part_ms'].items()}, d['config']['launch'])
except Exception as e:
    print('$tag failed', e); print(open('$OUT/b_$tag.err').read()[-2500:])
PY
done; done
# full-size N=2 with graph
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $OUT/b_full2.json 2> $OUT/b_full2.err
python -c "
import json
d=[json.loads(l) for l in open('$OUT/b_full2.json') if l.startswith('{')][-1]; print('full N=2', d['value'], d['ms_per_step'], d['part_ms'], d['config']['launch'])" || tail -20 $OUT/b_full2.err
