#!/bin/bash
set -u
OUT=gpurun_out/r3e
mkdir -p $OUT
FQ="SELECT sensor, value FROM flow WHERE value >= 10"
for S in 0 512; do
ARK_FP_SCOUT=$S ARK_FP_LB_DELAY=0 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none -k regex:filter_project_tile -s 4 -c 2 python scripts/prof_query.py "$FQ" 16777216 1000000 4 0 3 2>&1 | grep -E "dram__|lts__|gpu__time|filter_project_tile" | tee -a $OUT/ncu_scout$S.txt
done
