#!/bin/bash
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -q -k "bulk or exchange or golden" ) > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
for mb in ${2:-4 5 6}; do
  DCCM_SFC_MINB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $OUT/sfc_minb$mb.json 2>$OUT/err$mb.log
  python -c "import json; d=json.load(open('$OUT/sfc_minb$mb.json')); print('minb',$mb,d['part_ms'], d['value'])"
done
