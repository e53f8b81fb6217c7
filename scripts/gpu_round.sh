#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list.  Usage: gpurun -- bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
if [ "$3" != "lite" ]; then
( time timeout 900 python -m pytest tests -m gpu -q -s ) > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
fi
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
if [ "$3" != "lite" ]; then
( timeout 600 python bench.py --steps 10 --warmup 3 --unfused --no-e2e --no-cpu ) > $OUT/bench_unfused.json 2>> $OUT/bench.err
( timeout 600 python bench.py --steps 10 --warmup 3 --reference-order --no-e2e --no-cpu ) > $OUT/bench_exact.json 2>> $OUT/bench.err
for w in T341_0p25deg T106_1deg T42x64 T42; do ( timeout 300 python bench.py --steps 20 --warmup 3 --workload $w --no-e2e --no-cpu ) > $OUT/bench_$w.json 2>> $OUT/bench.err; done
fi
echo "bench exit $?" >> $OUT/bench.err
if [ "$2" = "full" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'remap_|vdiff|sfc_exchange' -s 24 -c 8 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'remap_|bulkflux|vdiff|sfc_exchange|ocn_' -c 60 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
tail -8 $OUT/pytest.log; for f in $OUT/bench_*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['part_ms'])"; done; cat $OUT/bench.json | cut -c1-3000; tail -3 $OUT/bench.err
