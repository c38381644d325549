#!/bin/bash
# round-2: the look-back warp scouts the tile `lead` tiles ahead
set -u
OUT=gpurun_out/r3d
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
ARK_FP_SCOUT=256 timeout 600 python -m pytest tests/test_sql_filter_gpu.py tests/test_sql_fuzz_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(emit\|gather\|finalize\)" >> $OUT/ab.log; }
for S in 64 128 256 512 1024 2048; do
run "scout$S delay0"   ARK_FP_SCOUT=$S ARK_FP_LB_DELAY=0 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
done
run "scout256 delay800" ARK_FP_SCOUT=256 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "scout512 delay300" ARK_FP_SCOUT=512 ARK_FP_LB_DELAY=300 timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "scout512 delay0 fixed-only" ARK_FP_SCOUT=512 ARK_FP_LB_DELAY=0 timeout 120 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
run "no scout" timeout 120 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
grep -E "^==|filter_project" $OUT/ab.log | paste - - | sed 's/filter_project_tma_kernel//'
