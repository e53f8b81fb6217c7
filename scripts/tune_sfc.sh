#!/bin/bash
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT
for mb in 3 4 5 6 8; do
  DCCM_SFC_MINB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $OUT/sfc_minb$mb.json 2>$OUT/err$mb.log
  python -c "import json; d=json.load(open('$OUT/sfc_minb$mb.json')); print('minb',$mb,d['part_ms'], d['value'])"
done
