"""Hash join (ark_sql_process_tables) and hash repartition on the device vs the oracle.
Mirrors JoinOperation (crates/arkflow-plugin/src/buffer/join.rs:62-132; the reference has no tests for it).
Join output order is unspecified: compared as multisets of rows."""
import numpy as np
import pyarrow as pa
import pytest

from arkflow_b200.arrow_ffi import DeviceBatch
from arkflow_b200.dist import NativeEngine
from arkflow_b200.processor import ArkError, SqlProcessor
from oracle.sql_oracle import sql_join
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def rows(rb):
    return sorted(map(repr, zip(*[c.to_pylist() for c in rb.columns]))) if rb.num_columns else []


def check_join(tables, query):
    want = sql_join(tables, query)
    p = SqlProcessor({"query": query})
    got_host = p.process_tables(tables)
    got_dev = p.process_tables_device({k: DeviceBatch.from_arrow(v) for k, v in tables.items()}).to_arrow()
    for got in (got_host, got_dev):
        assert got.schema.names == want.schema.names, (got.schema.names, want.schema.names)
        assert [f.type for f in got.schema] == [f.type for f in want.schema]
        assert got.num_rows == want.num_rows, (got.num_rows, want.num_rows)
        assert rows(got) == rows(want)
    return want


def test_join_buffer_example_query(gpu):
    # examples/join_buffer_example.yaml:25 — 10 decoded messages per input, same id on both sides
    a = pa.record_batch({"id": pa.array([1625000000000] * 10, pa.int64()), "value": pa.array([10] * 10, pa.int64()), "sensor": pa.array(["temp_1"] * 10)})
    b = pa.record_batch({"id": pa.array([1625000000000] * 10, pa.int64()), "value": pa.array([20] * 10, pa.int64()), "sensor": pa.array(["temp_2"] * 10)})
    out = check_join({"flow_input1": a, "flow_input2": b}, "SELECT * FROM flow_input1 join flow_input2 on (flow_input1.id = flow_input2.id)")
    assert out.num_rows == 100 and out.schema.names == ["id", "value", "sensor", "id", "value", "sensor"]


def test_config4_unique_build_keys_on_sensor(gpu):
    # build side has each key exactly once, probe keys uniform over them
    K = 5000
    build = synth_batch(K, seed=1, key_space=K)
    uniq = pa.array(["temp_%07d" % i for i in np.random.default_rng(0).permutation(K)])
    build = pa.record_batch({"timestamp": build.column("timestamp"), "value": build.column("value"), "sensor": uniq})
    probe = synth_batch(200_000, seed=2, key_space=K)
    out = check_join({"p": probe, "b": build}, "SELECT * FROM p JOIN b ON p.sensor = b.sensor")
    assert out.num_rows == 200_000


def test_join_int_key_duplicates_nulls_and_projection(gpu):
    rng = np.random.default_rng(4)
    n1, n2 = 3000, 2000
    a = pa.record_batch({"k": pa.array([None if rng.random() < 0.1 else int(x) for x in rng.integers(0, 300, n1)], pa.int64()),
                         "av": pa.array(rng.random(n1), pa.float64()), "astr": pa.array(["a%d" % i for i in range(n1)]),
                         "ab": pa.array([bool(i & 1) for i in range(n1)], pa.bool_())})
    b = pa.record_batch({"k": pa.array([None if rng.random() < 0.1 else int(x) for x in rng.integers(100, 500, n2)], pa.int64()),
                         "bv": pa.array(rng.integers(0, 9, n2), pa.int64(), mask=rng.random(n2) < 0.2)})
    check_join({"a": a, "b": b}, "SELECT * FROM a JOIN b ON a.k = b.k")
    check_join({"a": a, "b": b}, "SELECT a.astr, b.bv, a.k AS key, ab FROM a INNER JOIN b ON b.k = a.k")
    check_join({"a": a, "b": b}, "SELECT x.*, y.bv FROM a AS x JOIN b y ON x.k = y.k")


def test_outer_joins(gpu):
    """LEFT / RIGHT [OUTER] JOIN: unmatched rows of the preserved side survive once with NULLs on the other side
    (every type: Int64, Float64, Utf8, Boolean), duplicates and NULL keys included."""
    rng = np.random.default_rng(14)
    n1, n2 = 4000, 1500
    a = pa.record_batch({"k": pa.array([None if rng.random() < 0.1 else int(x) for x in rng.integers(0, 300, n1)], pa.int64()),
                         "av": pa.array(rng.random(n1), pa.float64()), "astr": pa.array(["a%d" % i for i in range(n1)]),
                         "ab": pa.array([bool(i & 1) for i in range(n1)], pa.bool_())})
    b = pa.record_batch({"k": pa.array([None if rng.random() < 0.1 else int(x) for x in rng.integers(200, 600, n2)], pa.int64()),
                         "bv": pa.array(rng.integers(0, 9, n2), pa.int64(), mask=rng.random(n2) < 0.2),
                         "bstr": pa.array(["b-%d" % (i % 37) for i in range(n2)]), "bb": pa.array([i % 3 == 0 for i in range(n2)], pa.bool_())})
    for q in ("SELECT * FROM a LEFT JOIN b ON a.k = b.k", "SELECT * FROM a RIGHT JOIN b ON a.k = b.k",
              "SELECT * FROM a LEFT OUTER JOIN b ON b.k = a.k", "SELECT a.astr, b.bstr, b.bb, a.k AS ak, b.k AS bk FROM a RIGHT OUTER JOIN b ON a.k = b.k",
              "SELECT * FROM b LEFT JOIN a ON a.k = b.k"):
        out = check_join({"a": a, "b": b}, q)
        assert out.num_rows >= (n1 if "LEFT" in q and "FROM a" in q else 0)
    # nothing matches: every preserved row once, the other side all NULL
    c = pa.record_batch({"k": pa.array([10_000 + i for i in range(40)], pa.int64()), "y": pa.array(["y%d" % i for i in range(40)])})
    out = check_join({"a": a, "c": c}, "SELECT * FROM a LEFT JOIN c ON a.k = c.k")
    assert out.num_rows == n1 and out.column("y").null_count == n1
    out = check_join({"a": a, "c": c}, "SELECT * FROM a RIGHT JOIN c ON a.k = c.k")
    assert out.num_rows == 40 and out.column("astr").null_count == 40


def test_shipped_temporary_right_join_shape(gpu):
    """examples/redis_temporary_example.yaml:29: `SELECT * FROM flow right join redis_table on (flow.sensor = redis_table.x)` —
    the batch joined against the rows a temporary store returned; store rows without a flow row survive with NULL flow columns."""
    flow = synth_batch(20_000, seed=3, key_space=40)
    redis_table = pa.record_batch({"x": pa.array(["temp_%07d" % i for i in range(30, 55)]), "meta": pa.array(range(25), pa.int64())})
    out = check_join({"flow": flow, "redis_table": redis_table}, "SELECT * FROM flow right join redis_table on (flow.sensor = redis_table.x)")
    assert out.schema.names == ["timestamp", "value", "sensor", "x", "meta"]
    assert out.column("sensor").null_count == 15  # store keys 40..54 have no flow row


def test_join_long_string_keys_and_no_matches(gpu):
    rng = np.random.default_rng(8)
    keys = ["k%d-%s" % (i, "z" * int(rng.integers(0, 40))) for i in range(400)]
    a = pa.record_batch({"s": pa.array([keys[int(i)] for i in rng.integers(0, 400, 5000)]), "x": pa.array(range(5000), pa.int64())})
    b = pa.record_batch({"s": pa.array([keys[int(i)] for i in rng.integers(0, 400, 700)]), "y": pa.array(range(700), pa.int64())})
    check_join({"a": a, "b": b}, "SELECT * FROM a JOIN b ON a.s = b.s")
    c = pa.record_batch({"s": pa.array(["nope%d" % i for i in range(50)]), "y": pa.array(range(50), pa.int64())})
    out = check_join({"a": a, "c": c}, "SELECT * FROM a JOIN c ON a.s = c.s")
    assert out.num_rows == 0


def test_single_table_query_through_tables_entry_point(gpu):
    rb = synth_batch(1000, key_space=5)
    out = check_join({"flow_input1": rb}, "SELECT sensor, value FROM flow_input1 WHERE value >= 10")
    assert out.num_rows > 0


def test_join_errors(gpu):
    a = pa.record_batch({"k": pa.array([1], pa.int64())})
    with pytest.raises(ArkError):
        SqlProcessor({"query": "SELECT * FROM a JOIN b ON a.k = b.k"}).process_tables({"a": a})
    with pytest.raises(ArkError) as e:
        SqlProcessor({"query": "SELECT * FROM a JOIN b ON a.k = b.nope"}).process_tables({"a": a, "b": a})
    assert e.value.kind == "Process"


def test_hash_partition_groups_keys_and_keeps_rows(gpu):
    rb = synth_batch(100_000, key_space=997)
    eng = NativeEngine("SELECT * FROM flow")
    out, counts = eng.hash_partition(DeviceBatch.from_arrow(rb), "sensor", 8)
    got = out.to_arrow()
    assert sum(counts) == rb.num_rows and got.num_rows == rb.num_rows
    assert rows(got) == rows(rb)
    starts = np.concatenate([[0], np.cumsum(counts)])
    keys = got.column("sensor").to_pylist()
    owner = {}
    for p in range(8):
        for k in set(keys[starts[p]:starts[p + 1]]):
            assert owner.setdefault(k, p) == p
    assert min(counts) > 0


@pytest.mark.parametrize("ranks", [2, 4])
def test_repartitioned_join_on_one_gpu(gpu, ranks):
    """Simulated multi-rank join: partition both sides per 'rank', route partition p to owner p, join locally."""
    K = 3000
    q = "SELECT * FROM p JOIN b ON p.sensor = b.sensor"
    p_shards = [synth_batch(20_000, row0=r * 20_000, seed=3, key_space=K) for r in range(ranks)]
    b_all = pa.record_batch({"sensor": pa.array(["temp_%07d" % i for i in range(K)]), "w": pa.array(range(K), pa.int64())})
    b_shards = [b_all.slice(r * (K // ranks), K // ranks if r < ranks - 1 else K - (ranks - 1) * (K // ranks)) for r in range(ranks)]
    eng = NativeEngine(q)

    def split(shards, col):
        parts = []
        for s in shards:
            out, counts = eng.hash_partition(DeviceBatch.from_arrow(s), col, ranks)
            rb = out.to_arrow()
            st = np.concatenate([[0], np.cumsum(counts)])
            parts.append([rb.slice(int(st[i]), int(counts[i])) for i in range(ranks)])
        return parts

    pp, bp = split(p_shards, "sensor"), split(b_shards, "sensor")
    got_rows = []
    for owner in range(ranks):
        lp = pa.Table.from_batches([pp[s][owner] for s in range(ranks)]).combine_chunks().to_batches()[0]
        lb = pa.Table.from_batches([bp[s][owner] for s in range(ranks)]).combine_chunks().to_batches()[0]
        out = eng.join({"p": DeviceBatch.from_arrow(lp), "b": DeviceBatch.from_arrow(lb)}).to_arrow()
        got_rows += list(map(repr, zip(*[c.to_pylist() for c in out.columns])))
    full_p = pa.Table.from_batches(p_shards).combine_chunks().to_batches()[0]
    want = sql_join({"p": full_p, "b": b_all}, q)
    assert sorted(got_rows) == rows(want)
