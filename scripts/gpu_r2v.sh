#!/bin/bash
set -u
OUT=gpurun_out/r2v
mkdir -p $OUT
timeout 900 python -m pytest tests/test_join_gpu.py tests/test_dist_gpu.py tests/test_sql_fuzz_gpu.py tests/test_buffers_gpu.py tests/test_sql_aggregate_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --e2e-steps 6 > $OUT/bench.json 2>$OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json").read().strip().split("\n") if l.startswith("{")][-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["pinned"]["value"])
for k in ("groupby","join","window"):
    v=d.get(k,{}); print(k, {kk: v.get(kk) for kk in ("value","ms_per_step","ms_per_window","verified","error")}, v.get("kernels_ms"))
PY
