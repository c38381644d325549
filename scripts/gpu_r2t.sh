#!/bin/bash
set -u
N=${1:-2}
bash scripts/gpu_r2s.sh $N
bash scripts/gpu_r2e.sh $N
