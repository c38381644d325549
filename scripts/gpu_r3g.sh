#!/bin/bash
set -u
OUT=gpurun_out/r3g
mkdir -p $OUT
one() { echo "== $1" | tee -a $OUT/conc.log; shift; env "$@" timeout 300 python bench.py --no-sharded --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.1f G  conc %.1f G (%.3f ms/step)'%(d['value']/1e9, d['concurrent_callers']['value']/1e9, d['concurrent_callers']['ms_per_step']))" | tee -a $OUT/conc.log; }
one "default"
one "no sampler" ARK_BENCH_NO_SAMPLER=1
one "default again"
one "no sampler, steps 16" ARK_BENCH_NO_SAMPLER=1 ARK_X=1
