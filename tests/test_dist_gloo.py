"""Host logic of the multi-GPU GROUP BY path (arkflow_b200/dist.py) under gloo, world_size 2, on CPU.

The CUDA compute steps are replaced by a CPU engine built on the oracle (allowed: tests/ may use
oracle/); what is exercised here is the all-to-all(v) exchange of partition-ordered batches
(fixed-width, Utf8 and validity buffers), partition ownership and the keyless special case.
"""
import os
import zlib

import numpy as np
import pyarrow as pa
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arkflow_b200.arrow_ffi import DeviceBatch
from arkflow_b200.dist import distributed_group_by, exchange_partitions
from oracle.sql_oracle import sql_process
from oracle.synth import synth_batch


class OracleEngine:
    """partial = GROUP BY with SUM/COUNT states, ordered by crc32(key) % n_parts; final = SUM of states."""

    def __init__(self, keyed=True):
        self.keyed = keyed

    def partial_aggregate(self, batch, n_parts):
        rb = batch.to_arrow()
        if self.keyed:
            st = sql_process(rb, "SELECT sensor, SUM(value) AS s, COUNT(*) AS c FROM flow GROUP BY sensor")
            keys = st.column("sensor").to_pylist()
            part = np.array([zlib.crc32((k or "").encode()) % n_parts for k in keys], dtype=np.int64)
            order = np.argsort(part, kind="stable")
            st = st.take(pa.array(order))
            rows = [int((part == p).sum()) for p in range(n_parts)]
        else:
            st = sql_process(rb, "SELECT SUM(value) AS __acc0, COUNT(*) AS __acc1 FROM flow")
            rows = [1] + [0] * (n_parts - 1)
        return DeviceBatch.from_arrow(st, device="cpu"), rows

    def final_aggregate(self, partial):
        rb = partial.to_arrow()
        if self.keyed:
            if rb.num_rows == 0:
                return DeviceBatch.from_arrow(pa.record_batch({"sensor": pa.array([], pa.utf8()), "sum": pa.array([], pa.int64()), "cnt": pa.array([], pa.int64())}), device="cpu")
            out = sql_process(rb, "SELECT sensor, SUM(s) AS sum, SUM(c) AS cnt FROM flow GROUP BY sensor")
        else:
            if rb.num_rows == 0:
                out = pa.record_batch({"sum": pa.array([None], pa.int64()), "cnt": pa.array([0], pa.int64())})
            else:
                out = sql_process(rb, "SELECT SUM(__acc0) AS sum, SUM(__acc1) AS cnt FROM flow")
        return DeviceBatch.from_arrow(out, device="cpu")


def _worker(rank, world, port, tmpdir, keyed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 20_000
        rb = synth_batch(n, row0=rank * n, seed=42, key_space=257)
        local = DeviceBatch.from_arrow(rb, device="cpu")
        # 1) raw exchange: send rows [0, a) to rank 0 and [a, n) to rank 1, Utf8 + int64 + validity
        a = 1234 + 1000 * rank
        vals = pa.array([None if i % 7 == 0 else i for i in range(n)], pa.int64())
        probe = pa.record_batch({"sensor": rb.column("sensor"), "v": vals, "b": pa.array([i % 3 == 0 for i in range(n)], pa.bool_())})
        got = exchange_partitions(DeviceBatch.from_arrow(probe, device="cpu"), [a, n - a], None).to_arrow()
        w = pa.ipc.new_file(os.path.join(tmpdir, f"xchg_{rank}.arrow"), got.schema)
        w.write_batch(got)
        w.close()
        # 2) distributed GROUP BY
        out = distributed_group_by(OracleEngine(keyed), local, None).to_arrow()
        w = pa.ipc.new_file(os.path.join(tmpdir, f"out_{rank}.arrow"), out.schema)
        w.write_batch(out)
        w.close()
    finally:
        dist.destroy_process_group()


def _read(path):
    t = pa.ipc.open_file(path).read_all().combine_chunks()
    b = t.to_batches()
    return b[0] if b else pa.RecordBatch.from_arrays([pa.array([], f.type) for f in t.schema], schema=t.schema)


@pytest.mark.parametrize("keyed", [True, False])
def test_group_by_two_ranks_gloo(tmp_path, keyed):
    world, port = 2, 29500 + (os.getpid() % 2000) + (0 if keyed else 1)
    mp.spawn(_worker, args=(world, port, str(tmp_path), keyed), nprocs=world, join=True)
    n = 20_000
    full = pa.Table.from_batches([synth_batch(n, row0=r * n, seed=42, key_space=257) for r in range(world)]).combine_chunks().to_batches()[0]
    outs = [_read(os.path.join(tmp_path, f"out_{r}.arrow")) for r in range(world)]
    if keyed:
        want = sql_process(full, "SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor")
        wd = {k: (s, c) for k, s, c in zip(*[col.to_pylist() for col in want.columns])}
        gd = {}
        for o in outs:
            for k, s, c in zip(*[col.to_pylist() for col in o.columns]):
                assert k not in gd, "a group was produced by two ranks"
                gd[k] = (s, c)
        assert gd == wd
        assert all(o.num_rows > 0 for o in outs), "both ranks should own some groups"
    else:
        want = sql_process(full, "SELECT SUM(value), COUNT(*) FROM flow")
        assert outs[1].num_rows == 0
        assert [c.to_pylist() for c in outs[0].columns] == [c.to_pylist() for c in want.columns]


def test_exchange_preserves_rows_gloo(tmp_path):
    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path), True), nprocs=world, join=True)
    n = 20_000
    probes = []
    for r in range(world):
        rb = synth_batch(n, row0=r * n, seed=42, key_space=257)
        probes.append(pa.record_batch({"sensor": rb.column("sensor"), "v": pa.array([None if i % 7 == 0 else i for i in range(n)], pa.int64()),
                                       "b": pa.array([i % 3 == 0 for i in range(n)], pa.bool_())}))
    cuts = [1234, 2234]
    want0 = pa.Table.from_batches([probes[0].slice(0, cuts[0]), probes[1].slice(0, cuts[1])]).combine_chunks().to_batches()[0]
    want1 = pa.Table.from_batches([probes[0].slice(cuts[0]), probes[1].slice(cuts[1])]).combine_chunks().to_batches()[0]
    for r, want in ((0, want0), (1, want1)):
        got = _read(os.path.join(tmp_path, f"xchg_{r}.arrow"))
        assert got.num_rows == want.num_rows
        for name in want.schema.names:
            assert got.column(name).to_pylist() == want.column(name).to_pylist(), name
