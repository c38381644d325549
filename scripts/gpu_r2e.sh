#!/bin/bash
# round-2 multi-GPU call (run with gpurun --gpus N): NCCL / peer-memory tests and the sharded bench at N ranks
set -u
N=${1:-2}
OUT=gpurun_out/r2e_n$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/run_dist_nccl.py > $OUT/dist.log 2>&1; echo "dist rc=$?" >> $OUT/dist.log
grep -E "OK|EXCHANGE|Error|error|assert" $OUT/dist.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 600 $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench.json").read().strip().split("\n") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "pinned", d["e2e"]["pinned"]["value"], "frac", d["roofline"]["frac"])
    for k in ("groupby","join","window"):
        v=d.get(k,{})
        print(k, {kk: v.get(kk) for kk in ("value","ms_per_step","ms_per_window","verified","error","roofline_frac")}, v.get("exchange"))
except Exception as e:
    print("parse failed", e)
PY
