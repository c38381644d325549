"""GROUP BY building blocks of the multi-GPU path on the device (C ABI): partial aggregate with hash
partitioning, final merge.  One GPU simulates R ranks: each "rank" partial-aggregates its slice,
partition p of every rank is routed to rank p by hand (what exchange_partitions does with NCCL), and
each owner merges.  The union must equal the oracle on the whole table, with disjoint owners."""
import os
import subprocess
import sys

import numpy as np
import pyarrow as pa
import pytest
import torch

from arkflow_b200.arrow_ffi import DeviceBatch
from arkflow_b200.dist import NativeEngine
from oracle.sql_oracle import sql_process
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def slice_rows(rb, starts):
    return [rb.slice(starts[i], starts[i + 1] - starts[i]) for i in range(len(starts) - 1)]


@pytest.mark.parametrize("ranks", [2, 4, 8])
@pytest.mark.parametrize("query,keys,floats", [
    ("SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor", ["sensor"], []),
    ("SELECT sensor, AVG(value), MIN(value), MAX(value), COUNT(value) FROM flow WHERE value >= 3 GROUP BY sensor", ["sensor"], ["avg(flow.value)"]),
])
def test_partial_exchange_final_on_one_gpu(gpu, ranks, query, keys, floats):
    n = 40_000
    shards = [synth_batch(n, row0=r * n, seed=42, key_space=1531) for r in range(ranks)]
    eng = NativeEngine(query)
    partials = []
    for r in range(ranks):
        pb, rows = eng.partial_aggregate(DeviceBatch.from_arrow(shards[r]), ranks)
        rb = pb.to_arrow()
        assert sum(rows) == rb.num_rows
        starts = np.concatenate([[0], np.cumsum(rows)]).tolist()
        partials.append(slice_rows(rb, starts))
    full = pa.Table.from_batches(shards).combine_chunks().to_batches()[0]
    want = sql_process(full, query)
    wd = {tuple(r[k] for k in keys): r for r in want.to_pylist()}
    seen = {}
    for owner in range(ranks):
        received = pa.Table.from_batches([partials[src][owner] for src in range(ranks)]).combine_chunks()
        if received.num_rows == 0:
            continue
        out = eng.final_aggregate(DeviceBatch.from_arrow(received.to_batches()[0])).to_arrow()
        assert out.schema.names == want.schema.names
        for row in out.to_pylist():
            k = tuple(row[c] for c in keys)
            assert k not in seen, f"group {k} owned by two ranks"
            seen[k] = row
    assert seen.keys() == wd.keys()
    for k, w in wd.items():
        for name, wv in w.items():
            gv = seen[k][name]
            if name in floats:
                assert abs(gv - wv) <= 1e-9 * max(1.0, abs(wv))
            else:
                assert gv == wv, (k, name, gv, wv)


def _exchange_contexts(ranks, region_bytes):
    from arkflow_b200.dist import ExchangeContext

    ctxs = [ExchangeContext(r, ranks, region_bytes) for r in range(ranks)]
    handles = [c.handle() for c in ctxs]
    for c in ctxs:
        c.connect(handles)
    return ctxs


@pytest.mark.parametrize("ranks", [1, 2, 5, 8])
@pytest.mark.parametrize("query,keys,floats", [
    ("SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor", ["sensor"], []),
    ("SELECT sensor, AVG(value), MIN(value), MAX(value), COUNT(value) FROM flow WHERE value >= 3 GROUP BY sensor", ["sensor"], ["avg(flow.value)"]),
    ("SELECT value, COUNT(*), MAX(timestamp) FROM flow GROUP BY value", ["value"], []),
    ("SELECT COUNT(*), SUM(value) FROM flow", [], []),
])
def test_device_side_exchange_on_one_gpu(gpu, ranks, query, keys, floats):
    """csrc/group_exchange.cu with every rank's context in one process on one GPU (peers resolve to the owners'
    own pointers; the push / flag / merge protocol is the one peers run over NVLink).  All ranks push, then all merge;
    several steps reuse the double-buffered regions.  The union over ranks must equal the oracle on the whole table."""
    ctxs = _exchange_contexts(ranks, 1 << 20)
    engines = [NativeEngine(query) for _ in range(ranks)]
    for step in range(4):  # ≥ 3 steps: both parities reused, acks exercised
        n = 30_000 + 1000 * step
        shards = [synth_batch(n, row0=(step * ranks + r) * n, seed=42 + step, key_space=1531) for r in range(ranks)]
        for r in range(ranks):
            engines[r].group_by_push(DeviceBatch.from_arrow(shards[r]), ctxs[r])
        outs = [engines[r].group_by_merge(ctxs[r]) for r in range(ranks)]
        assert all(o is not None for o in outs)
        full = pa.Table.from_batches(shards).combine_chunks().to_batches()[0]
        want = sql_process(full, query)
        wd = {tuple(r[k] for k in keys): r for r in want.to_pylist()}
        seen = {}
        for o in outs:
            rb = o.to_arrow()
            assert rb.schema.names == want.schema.names
            for row in rb.to_pylist():
                k = tuple(row[c] for c in keys)
                assert k not in seen, f"group {k} owned by two ranks"
                seen[k] = row
        assert seen.keys() == wd.keys()
        for k, w in wd.items():
            for name, wv in w.items():
                gv = seen[k][name]
                if name in floats:
                    assert abs(gv - wv) <= 1e-9 * max(1.0, abs(wv))
                else:
                    assert gv == wv, (step, k, name, gv, wv)
    for c in ctxs:
        c.close()


def test_device_side_exchange_long_keys_fall_back_on_every_rank(gpu):
    """Keys longer than 12 bytes cannot travel inline: EVERY rank's merge reports it (None), also ranks whose own
    keys were short, and the following step works again."""
    ranks = 3
    ctxs = _exchange_contexts(ranks, 1 << 18)
    q = "SELECT sensor, COUNT(*) FROM flow GROUP BY sensor"
    engines = [NativeEngine(q) for _ in range(ranks)]
    short = synth_batch(5000, seed=5, key_space=50)
    long_keys = pa.record_batch({"timestamp": short.column("timestamp"), "value": short.column("value"),
                                 "sensor": pa.array([f"a_rather_long_sensor_name_{i % 7}" for i in range(5000)])})
    for r in range(ranks):
        engines[r].group_by_push(DeviceBatch.from_arrow(long_keys if r == 1 else short), ctxs[r])
    assert [engines[r].group_by_merge(ctxs[r]) for r in range(ranks)] == [None] * ranks
    for r in range(ranks):
        engines[r].group_by_push(DeviceBatch.from_arrow(short), ctxs[r])
    outs = [engines[r].group_by_merge(ctxs[r]) for r in range(ranks)]
    total = sum(sum(o.to_arrow().column("count(*)").to_pylist()) for o in outs)
    assert total == ranks * 5000
    for c in ctxs:
        c.close()


def test_partial_states_of_one_key_land_in_one_partition(gpu):
    eng = NativeEngine("SELECT sensor, COUNT(*) FROM flow GROUP BY sensor")
    owner = {}
    for seed in (1, 2, 3):
        rb = synth_batch(30_000, seed=seed, key_space=97)
        pb, rows = eng.partial_aggregate(DeviceBatch.from_arrow(rb), 8)
        keys = pb.to_arrow().column(0).to_pylist()
        starts = np.concatenate([[0], np.cumsum(rows)])
        for p in range(8):
            for k in keys[starts[p]:starts[p + 1]]:
                assert owner.setdefault(k, p) == p
    assert len(set(owner.values())) > 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_group_by_two_gpus_nccl(gpu, tmp_path):
    script = os.path.join(os.path.dirname(__file__), "run_dist_nccl.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_OK" in r.stdout


def test_ipc_descriptor_pull_of_partition_slices(gpu):
    """csrc/ipc_exchange.cu on one GPU: each simulated rank exports its hash-partitioned batch as an IPC
    descriptor; owner p pulls rows [start_p, start_p + count_p) of every source with
    ark_ipc_concat_slices_device (same-process sources resolve to the exporter's own pointers, the slicing /
    concatenation / offset rebasing is the code path peers take).  Compared with slicing on the host."""
    import ctypes as C

    from arkflow_b200 import _lib as L
    from arkflow_b200.dist import _ipc_export
    from arkflow_b200.processor import _check

    ranks = 3
    eng = NativeEngine("SELECT * FROM p JOIN b ON p.sensor = b.sensor")
    parted, rows, hosts = [], [], []
    for r in range(ranks):
        rb = synth_batch(20_000 + 1000 * r, row0=r * 50_000, seed=11, key_space=977)
        nulls = pa.array([None if i % 97 == 0 else v for i, v in enumerate(rb.column("value").to_pylist())], pa.int64())
        rb = pa.record_batch({"timestamp": rb.column("timestamp"), "value": nulls, "sensor": rb.column("sensor"),
                              "flag": pa.array([i % 3 == 0 for i in range(rb.num_rows)])})
        pb, pr = eng.hash_partition(DeviceBatch.from_arrow(rb), "sensor", ranks)
        parted.append(pb); rows.append(pr); hosts.append(pb.to_arrow())
    blobs = [_ipc_export(pb) for pb in parted]
    assert all(b is not None for b in blobs)
    for owner in range(ranks):
        row0 = (C.c_int64 * ranks)(*[sum(rows[s][:owner]) for s in range(ranks)])
        nrows = (C.c_int64 * ranks)(*[rows[s][owner] for s in range(ranks)])
        keep = [(C.c_uint8 * len(b)).from_buffer_copy(b) for b in blobs]
        ptrs = (C.POINTER(C.c_uint8) * ranks)(*[C.cast(k, C.POINTER(C.c_uint8)) for k in keep])
        sizes = (C.c_int64 * ranks)(*[len(b) for b in blobs])
        out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
        _check(L.lib().ark_ipc_concat_slices_device(ranks, ptrs, sizes, row0, nrows, C.byref(out_dev), C.byref(out_sch)))
        got = DeviceBatch.adopt(out_dev, out_sch).to_arrow()
        want = pa.Table.from_batches([hosts[s].slice(row0[s], nrows[s]) for s in range(ranks)]).combine_chunks().to_batches()[0]
        assert got.schema.names == want.schema.names
        for name in want.schema.names:
            assert got.column(name).equals(want.column(name)), (owner, name)
