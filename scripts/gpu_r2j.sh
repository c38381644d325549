#!/bin/bash
# round-2: filter tile kernel — polling back-off / delayed look-back
set -u
OUT=gpurun_out/r2j
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for T in 256 512; do
for S in 0 100 300 1000; do
run "two-level dt$T sleep$S" ARK_FP_THREADS=$T ARK_FP_LB_SLEEP=$S timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
for D in 300 700 1500; do
run "two-level dt$T delay$D" ARK_FP_THREADS=$T ARK_FP_LB_DELAY=$D timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
run "two-level dt$T delay700 sleep300" ARK_FP_THREADS=$T ARK_FP_LB_DELAY=700 ARK_FP_LB_SLEEP=300 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "chain dt$T sleep300" ARK_FP_LB=1 ARK_FP_THREADS=$T ARK_FP_LB_SLEEP=300 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
grep -E "^==|filter_project" $OUT/ab.log | paste - - | sed 's/filter_project_tma_kernel//'
