"""torchrun entry: GROUP BY over N GPUs with the NCCL all-to-all, checked against the oracle on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pyarrow as pa
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from arkflow_b200 import _lib as L
    from arkflow_b200.arrow_ffi import DeviceBatch
    from arkflow_b200.dist import NativeEngine, distributed_group_by
    from arkflow_b200.processor import _check
    from oracle.sql_oracle import sql_process
    from oracle.synth import synth_batch

    _check(L.lib().ark_b200_init(local))
    n = 200_000
    query = "SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor"
    eng = NativeEngine(query)
    local_rb = synth_batch(n, row0=rank * n, seed=42, key_space=10_007)
    out = distributed_group_by(eng, DeviceBatch.from_arrow(local_rb)).to_arrow()
    rows = out.to_pylist()
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        full = pa.Table.from_batches([synth_batch(n, row0=r * n, seed=42, key_space=10_007) for r in range(world)]).combine_chunks().to_batches()[0]
        want = {r["sensor"]: r for r in sql_process(full, query).to_pylist()}
        got = {}
        for part in gathered:
            for r in part:
                assert r["sensor"] not in got
                got[r["sensor"]] = r
        assert got == want, "distributed GROUP BY differs from the oracle"
        print(f"GROUPBY_OK world={world} groups={len(got)} per-rank={[len(p) for p in gathered]}")

    # ---- distributed join: probe rows sharded by rank, build side (unique keys) sharded by rank ----
    from arkflow_b200.dist import distributed_join
    from oracle.sql_oracle import sql_join

    K = 5000
    jq = "SELECT * FROM p JOIN b ON p.sensor = b.sensor"
    jeng = NativeEngine(jq)
    probe = synth_batch(50_000, row0=rank * 50_000, seed=7, key_space=K)
    lo, hi = rank * K // world, (rank + 1) * K // world
    build = pa.record_batch({"sensor": pa.array(["temp_%07d" % i for i in range(lo, hi)]), "w": pa.array(range(lo, hi), pa.int64())})
    jout = distributed_join(jeng, {"p": DeviceBatch.from_arrow(probe), "b": DeviceBatch.from_arrow(build)}, {"p": "sensor", "b": "sensor"}).to_arrow()
    jrows = list(map(repr, zip(*[c.to_pylist() for c in jout.columns])))
    jg = [None] * world
    dist.all_gather_object(jg, jrows)
    if rank == 0:
        full_p = pa.Table.from_batches([synth_batch(50_000, row0=r * 50_000, seed=7, key_space=K) for r in range(world)]).combine_chunks().to_batches()[0]
        full_b = pa.record_batch({"sensor": pa.array(["temp_%07d" % i for i in range(K)]), "w": pa.array(range(K), pa.int64())})
        want_rows = sorted(map(repr, zip(*[c.to_pylist() for c in sql_join({"p": full_p, "b": full_b}, jq).columns])))
        assert sorted(sum(jg, [])) == want_rows, "distributed JOIN differs from the oracle"
        print(f"JOIN_OK world={world} rows={len(want_rows)} per-rank={[len(x) for x in jg]}")
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
