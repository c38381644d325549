#!/bin/bash
# final tree, N ranks: bench only
set -u
N=${1:-4}
OUT=gpurun_out/r3i_n$N
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus $N > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json").read().strip().split("\n") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "pinned", d["e2e"]["pinned"]["value"], "frac", d["roofline"]["frac"], "conc", d["concurrent_callers"]["value"])
for k in ("groupby","join","window"):
    v=d.get(k,{}); print(k, {kk: v.get(kk) for kk in ("value","ms_per_step","ms_per_window","verified","error")})
PY
