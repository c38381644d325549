#!/bin/bash
set -u
OUT=gpurun_out/r2q
mkdir -p $OUT
for K in 0 1 2; do for T in 8 12 16 24; do
echo "== copy kind $K threads $T" | tee -a $OUT/e2e.log
ARK_STAGE_COPY=$K ARK_STAGE_THREADS=$T ARK_STAGE_TRACE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-sharded --no-cpu-baseline --device-threads 1 --e2e-steps 12 2>$OUT/trace_${K}_$T.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('pageable %.3f pinned %.3f ratio %.3f'%(e['value']/1e9, e['pinned']['value']/1e9, e['pageable_over_pinned']))" | tee -a $OUT/e2e.log
grep "\[stage\] 201" $OUT/trace_${K}_$T.log | tail -2 | tee -a $OUT/e2e.log
done; done
