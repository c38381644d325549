#!/bin/bash
# round-2: filter tile kernel with the two-level look-back
set -u
OUT=gpurun_out/r2i
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
timeout 900 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for R in 40 48 56; do
run "two-level dt256 maxr$R" ARK_FP_MAXR=$R timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
run "two-level dt512 maxr40" ARK_FP_THREADS=512 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "two-level dt256 ticket" ARK_FP_TICKET=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "two-level dt256 stride1" ARK_FP_DESC_STRIDE=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "two-level dt256 stride2" ARK_FP_DESC_STRIDE=2 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "chain dt256" ARK_FP_LB=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "nolookback dt256" ARK_FP_DEBUG=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "two-level fixed-only dt256" timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
run "two-level helping forced" ARK_FP_DEBUG=4 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
cat $OUT/ab.log | grep -E "^==|filter_project"
