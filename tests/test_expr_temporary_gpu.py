"""expr::evaluate_expr on the device and the `sql` processor's temporary_list enrichment join
(SURVEY.md §8(f) rank 2).  The first block mirrors the reference's own tests
(crates/arkflow-plugin/src/expr/mod.rs:124-212); contents beyond what they pin are compared with
oracle/sql_oracle.py (evaluate_expr / sql_join)."""
import numpy as np
import pyarrow as pa
import pytest

from arkflow_b200.expr import ColumnarValue, EvaluateResult, Expr, Resource, Temporary, evaluate_expr
from arkflow_b200.processor import ArkError, MessageBatch, SqlProcessor
from oracle import sql_oracle
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def names_batch():
    return pa.record_batch({"name": pa.array(["Alice", "Bob", "Charlie"])})


# ---- the reference's tests ---------------------------------------------------------------------------
def test_scalar_literal_expression(gpu):
    # expr/mod.rs:131-146: ` 0.9` over an Int32 column batch → ColumnarValue::Scalar(Float64(0.9))
    rb = pa.record_batch({"a": pa.array([4, 230, 21], pa.int32())})
    cv = evaluate_expr(" 0.9", rb)
    assert cv.is_scalar() and cv.array.type == pa.float64() and cv.scalar_value() == 0.9


def test_string_expr(gpu):
    # expr/mod.rs:148-176
    r = Expr("Expr", "concat(name, ' is here')").evaluate_expr(names_batch())
    assert r.kind == "Vec" and r.value == ["Alice is here", "Bob is here", "Charlie is here"]
    r = Expr("Value", "test value").evaluate_expr(names_batch())
    assert r.kind == "Scalar" and r.value == "test value"


def test_evaluate_result_get(gpu):
    # expr/mod.rs:178-189
    s = EvaluateResult("Scalar", "test")
    assert s.get(0) == "test" and s.get(1) == "test"
    v = EvaluateResult("Vec", ["a", "b"])
    assert v.get(0) == "a" and v.get(1) == "b" and v.get(2) is None


def test_error_cases(gpu):
    # expr/mod.rs:191-211
    for bad in ("invalid sql", "1 + name"):
        with pytest.raises(ArkError) as e:
            Expr("Expr", bad).evaluate_expr(names_batch())
        assert e.value.kind == "Process" and e.value.message.startswith("Failed to evaluate expression: ")


# ---- beyond the reference's assertions: against the oracle ---------------------------------------------
@pytest.mark.parametrize("text", ["sensor", "value * 2 + 1", "value >= 10", "concat('k:', sensor)", "concat(sensor, '-', sensor, '!')",
                                  "CAST(value AS DOUBLE) / 4", "timestamp", "1 + 2", "'literal'", "concat('a', 'b')", "10 / 4", "NOT (value < 3)"])
def test_evaluate_expr_matches_oracle(gpu, text):
    rb = synth_batch(5000, key_space=37)
    want_scalar, want = sql_oracle.evaluate_expr(text, rb)
    cv = evaluate_expr(text, rb)
    assert cv.is_scalar() == want_scalar
    assert cv.array.type == want.type
    assert cv.array.equals(want), text


def test_evaluate_expr_nulls_and_empty(gpu):
    rb = pa.record_batch({"name": pa.array(["a", None, "ccc", None]), "x": pa.array([1, None, 3, 4], pa.int64())})
    cv = evaluate_expr("concat(name, '|', name)", rb)
    assert cv.array.to_pylist() == ["a|a", "|", "ccc|ccc", "|"]  # NULL arguments count as empty strings
    assert Expr("Expr", "name").evaluate_expr(rb).value == ["a", "ccc"]  # filter_map drops NULLs (expr/mod.rs:66-69)
    cv = evaluate_expr("x + 1", rb)
    assert cv.array.to_pylist() == [2, None, 4, 5]
    empty = rb.slice(0, 0)
    for text, typ in (("concat(name, 'z')", pa.utf8()), ("x * 2", pa.int64()), ("name", pa.utf8())):
        cv = evaluate_expr(text, empty)
        assert not cv.is_scalar() and len(cv.array) == 0 and cv.array.type == typ
    with pytest.raises(ArkError):
        Expr("Expr", "x").evaluate_expr(rb)  # not a string array: "Failed to evaluate expression"
    with pytest.raises(ArkError) as e:
        Expr("Expr", "1 + 1").evaluate_expr(rb)
    assert "Unsupported scalar type: Int64" in e.value.message
    with pytest.raises(ArkError):
        evaluate_expr("sum(x)", rb)


def test_concat_in_select_list(gpu):
    rb = synth_batch(20000, key_space=100)
    for q in ("SELECT value, concat('id-', sensor, '!') AS tag, sensor FROM flow WHERE value >= 10",
              "SELECT concat(sensor, sensor) FROM flow",
              "SELECT concat(sensor, '/', sensor), timestamp FROM flow WHERE value < 3 LIMIT 7"):
        want = sql_oracle.sql_process(rb, q)
        got = SqlProcessor({"query": q}).process(MessageBatch.new_arrow(rb)).batches[0].record_batch
        assert got.schema.names == want.schema.names
        assert got.equals(want), q
    with pytest.raises(ArkError) as e:  # Float64 arguments are not rendered on the device (Int64 / Boolean / Utf8 are)
        SqlProcessor({"query": "SELECT concat(value, 'x') FROM flow"}).process(MessageBatch.new_arrow(synth_batch(100, value_kind=1)))
    assert e.value.kind == "Unsupported"


# ---- temporary_list ------------------------------------------------------------------------------------
class DictTemporary(Temporary):
    """A test double for the reference's redis temporary (temporary/redis.rs:60-120): key → row."""

    def __init__(self, rows: dict):
        self.rows, self.seen = rows, []

    def get(self, keys):
        cv = keys[0]
        ks = [cv.scalar_value()] if cv.is_scalar() else [k for k in cv.array.to_pylist() if k is not None]
        self.seen.append((cv.kind, len(ks)))
        hit = sorted({k for k in ks if k in self.rows})
        if not hit:
            return None
        return MessageBatch.new_arrow(pa.record_batch({"x": pa.array(hit), "weight": pa.array([self.rows[k] for k in hit], pa.int64())}))


def test_temporary_list_enrichment_join(gpu):
    rb = synth_batch(3000, key_space=40)
    keys = sorted(set(rb.column("sensor").to_pylist()))
    store = DictTemporary({k: i * 10 for i, k in enumerate(keys) if i % 2 == 0})
    q = "SELECT flow.sensor, value, weight FROM flow JOIN t ON flow.sensor = t.x"
    p = SqlProcessor({"query": q, "temporary_list": [{"name": "kv", "table_name": "t", "key": {"type": "expr", "expr": "sensor"}}]},
                     Resource(temporary={"kv": store}))
    got = p.process(MessageBatch.new_arrow(rb)).batches[0].record_batch
    assert store.seen == [("Array", 3000)]
    table = store.get([ColumnarValue("Array", rb.column("sensor"))]).record_batch
    want = sql_oracle.sql_join({"flow": rb, "t": table}, q)
    key = lambda b: sorted(zip(*[b.column(i).to_pylist() for i in range(b.num_columns)]))
    assert got.schema.names == want.schema.names and key(got) == key(want)
    assert got.num_rows == sum(1 for s in rb.column("sensor").to_pylist() if s in store.rows)


def test_temporary_list_value_key_and_missing_temporary(gpu):
    rb = synth_batch(100, key_space=5)
    store = DictTemporary({"test": 7})
    p = SqlProcessor({"query": "SELECT value FROM flow WHERE value >= 10",
                      "temporary_list": [{"name": "kv", "table_name": "t", "key": {"type": "value", "value": "test"}}]},
                     Resource(temporary={"kv": store}))
    got = p.process(MessageBatch.new_arrow(rb)).batches[0].record_batch
    assert store.seen == [("Scalar", 1)]  # Expr::Value → ColumnarValue::Scalar(Utf8) (sql.rs:164-166)
    assert got.equals(sql_oracle.sql_process(rb, "SELECT value FROM flow WHERE value >= 10"))
    with pytest.raises(ArkError) as e:  # sql.rs:73-79
        SqlProcessor({"query": "SELECT * FROM flow", "temporary_list": [{"name": "nope", "table_name": "t", "key": {"type": "value", "value": "k"}}]},
                     Resource())
    assert e.value.kind == "Process" and e.value.message == "Temporary nope not found"
    with pytest.raises(ArkError) as e:
        bad = SqlProcessor({"query": "SELECT * FROM flow", "temporary_list": [{"name": "kv", "table_name": "t", "key": {"type": "expr", "expr": "no_such_col"}}]},
                           Resource(temporary={"kv": store}))
        bad.process(MessageBatch.new_arrow(rb))
    assert e.value.message.startswith("Evaluate expression failed: ")


def test_cast_to_string_and_numeric_concat_arguments(gpu):
    # examples/mqtt_example.yaml: SELECT * ,cast(value as string) as tx FROM flow WHERE value > 10
    rb = synth_batch(30_000, key_space=40)
    q = "SELECT * ,cast(value as string) as tx FROM flow WHERE value > 10"
    want = sql_oracle.sql_process(rb, q)
    got = SqlProcessor({"query": q}).process(MessageBatch.new_arrow(rb)).batches[0].record_batch
    assert got.schema.names == want.schema.names == ["timestamp", "value", "sensor", "tx"]
    assert got.equals(want)
    mixed = pa.record_batch({"value": pa.array([5, None, -12, 30, -(2**63), 2**63 - 1], pa.int64()),
                             "flag": pa.array([True, False, None, True, False, None]),
                             "s": pa.array(["a", "b", None, "d", "", "ünï"])})
    for q in ("SELECT cast(value as string) AS tx, CAST(flag AS STRING) AS f, value FROM flow",
              "SELECT concat(s, '#', value, '/', flag) AS c FROM flow",
              "SELECT concat(value, value) FROM flow WHERE value IS NOT NULL"):
        want = sql_oracle.sql_process(mixed, q)
        got = SqlProcessor({"query": q}).process(MessageBatch.new_arrow(mixed)).batches[0].record_batch
        assert got.schema.names == want.schema.names
        for name in want.schema.names:
            assert got.column(name).to_pylist() == want.column(name).to_pylist(), (q, name)
    cv = evaluate_expr("cast(value as string)", mixed)
    assert cv.array.to_pylist() == ["5", None, "-12", "30", str(-(2**63)), str(2**63 - 1)]
