#!/bin/bash
# round-2: filter tile kernel — 32-register variant (7 CTAs per SM)
set -u
OUT=gpurun_out/r2m
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for T in 256 512; do
for D in 0 500 1000; do
run "two-level dt$T maxr32 delay$D" ARK_FP_THREADS=$T ARK_FP_MAXR=32 ARK_FP_LB_DELAY=$D timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
run "nolookback dt$T maxr32" ARK_FP_DEBUG=1 ARK_FP_MAXR=32 ARK_FP_THREADS=$T timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "fixed-only dt$T maxr32" ARK_FP_THREADS=$T ARK_FP_MAXR=32 timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
done
run "fixed-only dt256 maxr40 delay1000" ARK_FP_LB_DELAY=1000 timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
grep -E "^==|filter_project" $OUT/ab.log | paste - - | sed 's/filter_project_tma_kernel//'
ARK_FP_MAXR=32 timeout 900 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
