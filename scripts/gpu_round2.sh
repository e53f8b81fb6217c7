#!/bin/bash
# Round-2 GPU-box visit (1 GPU): parity tests, bench (both arms), ncu.   gpurun --timeout 1500 -- bash scripts/gpu_round2.sh TAG [full]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt; lscpu | grep -i "model name\|numa\|socket" >> $OUT/host.txt
( time timeout 900 python -m pytest tests -m gpu -q -s ) > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
( timeout 300 python bench.py --steps 10 --warmup 3 --fast --no-e2e --no-cpu --no-others --no-parity ) > $OUT/bench_fast.json 2>> $OUT/bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
if [ "$2" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'remap_|vdiff|sfc_exchange' -s 27 -c 9 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others --no-parity > $OUT/ncu_full.log 2>&1
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'remap_|bulkflux|vdiff|sfc_exchange|ocn_' -c 72 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others --no-parity > $OUT/ncu_bench.log 2>&1
tail -6 $OUT/pytest.log
python - <<PY
import json
for f in ("bench.json", "bench_fast.json", "bench_ref.json"):
    try:
        d = [json.loads(l) for l in open("$OUT/" + f) if l.startswith("{")][-1]
        print(f, d.get("value"), d.get("ms_per_step"), d.get("part_ms"), d.get("output_hash"))
        for k in ("parity", "e2e", "e2e_dropin", "full_grid_check"):
            if k in d: print("  ", k, json.dumps(d[k])[:600])
        for w, v in d.get("other_workloads", {}).items():
            print("  ", w, v.get("value"), v.get("parity", {}).get("bitwise"), v.get("parity", {}).get("max_rel"), v.get("error"))
        if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 $OUT/bench.err
