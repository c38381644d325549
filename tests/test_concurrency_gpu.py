"""Re-entrancy of the C ABI: the reference drives `thread_num` pipeline workers through one processor object
(crates/arkflow-core/src/stream/mod.rs:117-126), so several host threads call process() on the same handles at
once.  Mixed queries run concurrently here (different plans, table-size hints, stream leases, pooled blocks) and
every result is compared with the oracle."""
import json
from concurrent.futures import ThreadPoolExecutor

import pyarrow as pa
import pytest

from arkflow_b200.processor import JsonToArrowProcessor, MessageBatch, SqlProcessor
from oracle.json_oracle import json_to_arrow
from oracle.sql_oracle import sql_join, sql_process
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def rows(rb):
    return sorted(map(repr, zip(*[c.to_pylist() for c in rb.columns])))


def test_mixed_queries_from_many_threads(gpu):
    batches = [synth_batch(60_000 + 1000 * i, row0=100_000 * i, seed=50 + i, key_space=ks) for i, ks in enumerate((3, 50, 700, 40_000))]
    queries = [
        "SELECT sensor, value FROM flow WHERE value >= 10",
        "SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor",
        "SELECT sensor, MIN(value), MAX(value), COUNT(value) FROM flow WHERE value < 15 GROUP BY sensor",
        "SELECT timestamp, value * 2 + 1 AS v FROM flow WHERE value % 3 = 0",
        "SELECT COUNT(*), SUM(value) FROM flow",
    ]
    procs = [SqlProcessor({"query": q}) for q in queries]
    want = {(qi, bi): sql_process(b, q) for qi, q in enumerate(queries) for bi, b in enumerate(batches)}
    build = pa.record_batch({"sensor": pa.array(["temp_%07d" % i for i in range(0, 700, 2)]), "w": pa.array(range(0, 700, 2), pa.int64())})
    jq = "SELECT p.sensor, value, w FROM p JOIN b ON p.sensor = b.sensor"
    jproc = SqlProcessor({"query": jq})
    jwant = rows(sql_join({"p": batches[2], "b": build}, jq))
    payloads = MessageBatch.new_binary([json.dumps({"timestamp": i, "value": i % 20, "sensor": f"temp_{i % 5}"}).encode() for i in range(20_000)])
    dec = JsonToArrowProcessor({})
    dwant = json_to_arrow(payloads.record_batch)

    def work(t):
        out = []
        for rep in range(6):
            for qi in range(len(queries)):
                bi = (t + rep + qi) % len(batches)
                got = procs[qi].process(MessageBatch.new_arrow(batches[bi])).batches[0].record_batch
                w = want[(qi, bi)]
                assert got.schema.names == w.schema.names
                if "GROUP BY" in queries[qi] or qi == 4:
                    assert rows(got) == rows(w), (t, rep, qi, bi)
                else:
                    assert got.equals(w), (t, rep, qi, bi)
            assert rows(jproc.process_tables({"p": batches[2], "b": build})) == jwant
            assert dec.process(payloads).batches[0].record_batch.equals(dwant)
            out.append(rep)
        return len(out)

    with ThreadPoolExecutor(max_workers=6) as pool:
        assert list(pool.map(work, range(6))) == [6] * 6
