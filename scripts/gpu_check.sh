#!/bin/bash
# quick validation: smoke + gpu tests + default bench
OUT=gpurun_out/${1:-chk}; mkdir -p $OUT
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
( time timeout 600 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" | tee -a $OUT/bench.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"
tail -3 $OUT/smoke.log; tail -4 $OUT/pytest.log; cut -c1-1500 $OUT/bench.json; tail -3 $OUT/bench.err; cut -c1-400 $OUT/bench_ref.json
