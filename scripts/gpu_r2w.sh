#!/bin/bash
# round-2: persistent ring kernel (store of tile k after the count of tile k+1)
set -u
OUT=gpurun_out/r2y
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
ARK_FP_IMPL=3 ARK_FP_THREADS=128 timeout 600 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
run "ring"              ARK_FP_IMPL=3 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "ring dt128 nolookback" ARK_FP_IMPL=3 ARK_FP_THREADS=128 ARK_FP_DEBUG=1 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "ring nolookback"   ARK_FP_IMPL=3 ARK_FP_DEBUG=1 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "ring dt128"        ARK_FP_IMPL=3 ARK_FP_THREADS=128 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "ring ts,value,sensor" ARK_FP_IMPL=3 timeout 120 python scripts/prof_query.py "SELECT timestamp, value, sensor FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
run "tile (default)"    timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
grep -E "^==|filter_project" $OUT/ab.log | paste - - | sed 's/filter_project_tma_kernel//'
