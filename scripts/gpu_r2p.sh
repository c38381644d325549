#!/bin/bash
set -u
OUT=gpurun_out/r2p
mkdir -p $OUT
nproc; lscpu | grep -E "Model name|NUMA|Socket|Thread" 
timeout 600 python scripts/prof_staging.py 2>&1 | tee $OUT/memcpy.log
for ET in 1 3; do
echo "== e2e threads $ET" | tee -a $OUT/trace.log
ARK_STAGE_TRACE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-sharded --no-cpu-baseline --device-threads 1 --e2e-threads $ET --e2e-steps 6 2>>$OUT/trace.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(json.dumps(d['e2e']))" | tee -a $OUT/trace.log
done
grep "\[stage\]" $OUT/trace.log | tail -30
