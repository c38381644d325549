#!/bin/bash
# multi-GPU diagnostics: phases of the distributed join; staging chunk size / copy kind under host contention
set -u
N=${1:-2}
OUT=gpurun_out/r2r_n$N
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 scripts/prof_dist_join.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee $OUT/join_phases.log
for CFG in "4096 2" "4096 0" "1024 2" "1024 0" "256 0" "256 2"; do
set -- $CFG
echo "== chunk $1 KB copy kind $2" | tee -a $OUT/e2e.log
ARK_STAGE_CHUNK_KB=$1 ARK_STAGE_COPY=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus $N --steps 8 --warmup 3 --no-sharded --no-cpu-baseline --device-threads 1 --e2e-steps 12 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().strip().split('\n') if l.startswith('{')][-1]); e=d['e2e']; print('pageable %.3f pinned %.3f ratio %.3f'%(e['value']/1e9, e['pinned']['value']/1e9, e['pageable_over_pinned']))" | tee -a $OUT/e2e.log
done
