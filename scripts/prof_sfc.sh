#!/bin/bash
# ncu --set full of the staged fused surface kernel + launch list of one whole exchange.  Usage: gpurun -- bash scripts/prof_sfc.sh TAG
OUT=gpurun_out/${1:-prof}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sfc_exchange_staged|sfc_exchange_kernel' -s 3 -c 1 \
    -o $OUT/sfc_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'remap_|bulkflux|vdiff|sfc_exchange|ocn_' -s 27 -c 27 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
grep -c "gpu__time_duration" $OUT/launches.csv
