#!/bin/bash
# compute-sanitizer over the device code, on one GPU:  gpurun --timeout 1500 -- bash scripts/gpu_sanitize.sh [tag]
# smoke() = one small exchange (T21 <-> Pl42-like grids) through every kernel of the path, unfused and fused (staged:
# TMA bulk copies + mbarrier + shared-memory stencil records -> racecheck / synccheck are the interesting tools).
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for tool in memcheck racecheck synccheck initcheck; do
  ( time timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke ) > $OUT/$tool.log 2>&1
  echo "$tool exit $?" | tee -a $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" $OUT/$tool.log | tail -3
done
# the late tests (second-order stencils: three source rows per tile) under memcheck as well
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_late_gpu.py -m gpu -x -q ) > $OUT/memcheck_late.log 2>&1
echo "memcheck late tests exit $?" | tee -a $OUT/memcheck_late.log
# round 2: the forward solve's redo path, the slab pipeline (row-range launches on two streams) and the portable exp / log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_zz_late_gpu.py -m gpu -x -q \
    -k "redo or slab or portable or extreme or reference_order" ) > $OUT/memcheck_round2.log 2>&1
echo "memcheck round-2 tests exit $?" | tee -a $OUT/memcheck_round2.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_zz_late_gpu.py -m gpu -x -q -k "slab" ) > $OUT/racecheck_slab.log 2>&1
echo "racecheck slab pipeline exit $?" | tee -a $OUT/racecheck_slab.log

