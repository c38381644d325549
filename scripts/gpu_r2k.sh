#!/bin/bash
# round-2: filter tile kernel without the start-up barrier
set -u
OUT=gpurun_out/r2k
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
timeout 900 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for T in 256 512; do
for R in 40 48 56; do
run "two-level dt$T maxr$R" ARK_FP_THREADS=$T ARK_FP_MAXR=$R timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
run "two-level dt$T delay500" ARK_FP_THREADS=$T ARK_FP_LB_DELAY=500 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "chain dt$T" ARK_FP_LB=1 ARK_FP_THREADS=$T timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "ticket dt$T" ARK_FP_TICKET=1 ARK_FP_THREADS=$T timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "nolookback dt$T" ARK_FP_DEBUG=1 ARK_FP_THREADS=$T timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "fixed-only dt$T" ARK_FP_THREADS=$T timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
done
run "r1 512" ARK_FP_IMPL=1 ARK_FP_THREADS=512 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
grep -E "^==|filter_project" $OUT/ab.log | paste - - | sed 's/filter_project_tma_kernel//'
