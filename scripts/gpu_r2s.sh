#!/bin/bash
set -u
N=${1:-2}
OUT=gpurun_out/r2s_n$N
mkdir -p $OUT
ARK_DIST_TRACE=1 ARK_KERNEL_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 scripts/prof_dist_join.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee $OUT/join_phases.log | tail -40
