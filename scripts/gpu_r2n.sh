#!/bin/bash
# round-2: blocked-row kernel (round 1) with the two-level look-back and a ticket
set -u
OUT=gpurun_out/r2n
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for T in 256 512; do
for LB in 1 2; do
run "r1 dt$T lb$LB" ARK_FP_IMPL=1 ARK_FP_THREADS=$T ARK_FP_LB=$LB timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "r1 dt$T lb$LB ticket" ARK_FP_IMPL=1 ARK_FP_THREADS=$T ARK_FP_LB=$LB ARK_FP_TICKET=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
run "r1 dt$T lb2 delay500" ARK_FP_IMPL=1 ARK_FP_THREADS=$T ARK_FP_LB_DELAY=500 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "r1 dt$T lb2 fixed-only" ARK_FP_IMPL=1 ARK_FP_THREADS=$T timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
done
grep -E "^==|filter_project" $OUT/ab.log | paste - - | sed 's/filter_project_tma_kernel//'
ARK_FP_IMPL=1 ARK_FP_THREADS=512 ARK_FP_TICKET=1 timeout 900 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
