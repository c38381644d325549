#!/bin/bash
# round-2: filter tile kernel — register cap x tile size sweep, then the filter tests under the 2048-row configuration
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for T in 256 512; do for R in 40 48 56; do
run "filter tile dt$T maxr$R" ARK_FP_THREADS=$T ARK_FP_MAXR=$R timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done; done
run "filter tile dt512 maxr40 ticket" ARK_FP_THREADS=512 ARK_FP_MAXR=40 ARK_FP_TICKET=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter tile dt512 maxr40 nolookback" ARK_FP_THREADS=512 ARK_FP_MAXR=40 ARK_FP_DEBUG=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter fixed-only dt512 maxr40" ARK_FP_THREADS=512 ARK_FP_MAXR=40 timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
run "filter fixed-only dt512 maxr56" ARK_FP_THREADS=512 ARK_FP_MAXR=56 timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
cat $OUT/ab.log
for R in 40 56; do
ARK_FP_THREADS=512 ARK_FP_MAXR=$R timeout 900 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest_dt512_$R.log 2>&1; echo "rc=$?" >> $OUT/pytest_dt512_$R.log; tail -3 $OUT/pytest_dt512_$R.log
done
