#!/bin/bash
# round-2 fourth GPU call: look-back warp + helping filter kernel, e2e staging knobs, full tests
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -6 $OUT/pytest.log
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
run() { echo "== $1" >> $OUT/ab.log; shift; env "$@" 2>&1 | grep -v "^agg_\(init\|compact\|emit\|gather\|finalize\)" >> $OUT/ab.log; }
run "filter tile lbwarp minb5"     ARK_FP_IMPL=2 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter tile lbwarp minb4"     ARK_FP_IMPL=2 ARK_FP_MINB=4 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter tile lbwarp minb6"     ARK_FP_IMPL=2 ARK_FP_MINB=6 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter tile lbwarp ticket"    ARK_FP_IMPL=2 ARK_FP_TICKET=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter tile nolookback"       ARK_FP_IMPL=2 ARK_FP_DEBUG=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter tile helping forced"   ARK_FP_IMPL=2 ARK_FP_DEBUG=4 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter r1 kernel"             ARK_FP_IMPL=1 timeout 300 python scripts/prof_query.py "$FQ" 16777216 1000000 20 0 6
run "filter fixed-only tile"       ARK_FP_IMPL=2 timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
run "filter fixed-only tile minb4" ARK_FP_IMPL=2 ARK_FP_MINB=4 timeout 300 python scripts/prof_query.py "SELECT timestamp, value FROM flow WHERE value >= 10" 16777216 1000000 20 0 6
cat $OUT/ab.log
for T in 4 8 16; do
  echo "== e2e ARK_STAGE_THREADS=$T" >> $OUT/e2e.log
  ARK_STAGE_THREADS=$T timeout 600 python bench.py --steps 8 --warmup 3 --no-sharded --no-cpu-baseline --device-threads 1 2>>$OUT/e2e.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(json.dumps({'value':d['value'],'frac':d['roofline']['frac'],'e2e':d['e2e']}))" >> $OUT/e2e.log
done
echo "== e2e threads 5" >> $OUT/e2e.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-sharded --no-cpu-baseline --device-threads 1 --e2e-threads 5 2>>$OUT/e2e.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(json.dumps({'e2e':d['e2e']}))" >> $OUT/e2e.log
cat $OUT/e2e.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_project_tile -s 6 -c 2 -o $OUT/fp_tile python scripts/prof_query.py "$FQ" 16777216 1000000 4 0 3 > $OUT/ncu_fp.log 2>&1
ls -la $OUT
