"""CPU-side checks of the C ABI: the library loads, exports every symbol include/arkflow_b200.h
declares, and the construction-time behaviour of the processors (no CUDA needed: SQL is parsed at
build time like the reference, sql.rs:91-98)."""
import pytest

from arkflow_b200 import _lib as L
from arkflow_b200.processor import ArkError, SqlProcessor, build_processor, init, register_processor_builder, split_batch
from golden_util import load_pins


def test_library_exports_every_declared_symbol(lib):
    declared = set(L.declared_symbols())
    assert declared, "no symbols parsed from include/arkflow_b200.h"
    assert declared == set(L.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)


def test_version_and_thread_local_error(lib):
    assert b"sm_100a" in lib.ark_b200_version()


@pytest.mark.parametrize("pin", [p for p in load_pins() if p["expect"]["kind"] == "ConstructError"], ids=lambda p: p["id"])
def test_construction_errors_match_reference(lib, pin):
    with pytest.raises(ArkError) as e:
        SqlProcessor(None if pin["query"] is None else {"query": pin["query"]})
    assert e.value.kind == pin["expect"]["error_kind"]
    assert e.value.message.startswith(pin["expect"]["prefix"])


def test_config_shape_errors(lib):
    with pytest.raises(ArkError) as e:
        SqlProcessor({"table_name": "t"})  # serde: missing field `query` → Error::Serialization via `?` (sql.rs:240)
    assert e.value.kind == "Serialization"
    with pytest.raises(ArkError) as e:
        SqlProcessor({"query": 5})
    assert e.value.kind == "Serialization"
    with pytest.raises(ArkError) as e:
        SqlProcessor({"query": "SELECT 1 FROM flow", "temporary_list": [{"name": "redis1", "table_name": "t", "key": {"type": "value", "value": "x"}}]})
    assert e.value.kind == "Process" and "Temporary redis1 not found" in e.value.message  # sql.rs:73-78


@pytest.mark.parametrize("q", [
    "SELECT * FROM flow", "select Sensor, VALUE from FLOW where value>=10", "SELECT sum(value),avg(value) ,111 as x FROM flow  group by sensor",
    "SELECT *,cast( __value__  as string) as y FROM flow ", "SELECT count(*) FROM flow WHERE value >= 10 group by sensor",
    "SELECT * FROM flow_input1 join flow_input2 on (flow_input1.id = flow_input2.id)",
    "SELECT a.x, b.y FROM t1 AS a INNER JOIN t2 b ON a.k = b.k", "SELECT * FROM a LEFT JOIN b ON a.k = b.k",
    "SELECT * FROM flow right join redis_table on (flow.sensor = redis_table.x)", "SELECT * FROM a LEFT OUTER JOIN b ON a.k = b.k", "SELECT \"Weird Name\", -value, value % 3 FROM flow WHERE NOT (value IS NULL) LIMIT 5;",
])
def test_accepted_sql(lib, q):
    SqlProcessor({"query": q})


@pytest.mark.parametrize("q,kind", [
    ("SELEC * FROM flow", "Process"), ("SELECT FROM flow", "Process"), ("SELECT * FROM", "Process"), ("SELECT * FROM flow WHERE", "Process"),
    ("SELECT (1 FROM flow", "Process"), ("DROP TABLE flow", "Process"), ("INSERT INTO flow VALUES (1)", "Process"), ("", "Process"),
    ("SELECT * FROM flow ORDER BY value", "Unsupported"), ("SELECT DISTINCT sensor FROM flow", "Unsupported"),
    ("SELECT * FROM (SELECT * FROM flow) t", "Unsupported"), ("SELECT CASE WHEN value > 1 THEN 1 END FROM flow", "Unsupported"),
    ("SELECT * FROM a FULL JOIN b ON a.k = b.k", "Unsupported"), ("SELECT * FROM a CROSS JOIN b", "Unsupported"),
    ("SELECT * FROM a LEFT b ON a.k = b.k", "Process"),
])
def test_rejected_sql(lib, q, kind):
    with pytest.raises(ArkError) as e:
        SqlProcessor({"query": q})
    assert e.value.kind == kind
    if kind == "Process":
        assert e.value.message.startswith("SQL query error")


def test_registry_mirrors_reference(lib):
    init()
    init()  # idempotent here; the reference's init() errors on duplicates via register_processor_builder
    with pytest.raises(ArkError) as e:
        register_processor_builder("sql", lambda n, c: None)  # core/processor/mod.rs:121-126
    assert e.value.kind == "Config"
    p = build_processor({"type": "sql", "query": "SELECT * FROM flow"})
    assert isinstance(p, SqlProcessor)
    with pytest.raises(ArkError):
        build_processor({"type": "nope"})


def test_split_batch_matches_reference_rules():
    # core/lib.rs:432-458
    import pyarrow as pa

    rb = pa.record_batch({"x": pa.array(range(100_000), pa.int64())})
    assert len(split_batch(rb.slice(0, 8192), 4)) == 1
    parts = split_batch(rb, 4)            # 4 * 8192 < 100000 → ceil(100000/4) rows each
    assert [p.num_rows for p in parts] == [25_000] * 4
    parts = split_batch(rb, 64)           # 64 * 8192 >= 100000 → 8192-row chunks
    assert [p.num_rows for p in parts][:2] == [8192, 8192] and sum(p.num_rows for p in parts) == 100_000
    assert len(split_batch(rb, 0)) == 1 or sum(p.num_rows for p in split_batch(rb, 0)) == 100_000


def test_sliding_window_and_batch_builders_validate_config(lib):
    # buffer/sliding_window.rs:248-270, 318-394 and processor/batch.rs:135-139 — all decided before any CUDA call
    from arkflow_b200.buffer import SlidingWindow, build_buffer
    from arkflow_b200.processor import BatchProcessor

    for cfg, text in (({"window_size": 0, "interval": "1s", "slide_size": 5}, "window_size must be greater than 0"),
                      ({"window_size": 10, "interval": "1s", "slide_size": 0}, "slide_size must be greater than 0"),
                      ({"window_size": 5, "interval": "1s", "slide_size": 10}, "window_size must be greater than slide_size")):
        with pytest.raises(ArkError) as e:
            build_buffer({"type": "sliding_window", **cfg})
        assert e.value.kind == "Config" and text in e.value.message
    with pytest.raises(ArkError) as e:
        SlidingWindow(None)
    assert e.value.kind == "Config"
    with pytest.raises(ArkError) as e:
        SlidingWindow({"window_size": 3, "interval": "soon", "slide_size": 1})
    assert e.value.kind == "Serialization"
    with pytest.raises(ArkError) as e:
        BatchProcessor(None)
    assert e.value.kind == "Config" and e.value.message == "Batch processor configuration is missing"
    with pytest.raises(ArkError) as e:
        BatchProcessor({"count": 2})
    assert e.value.kind == "Serialization"
    BatchProcessor({"count": 2, "timeout_ms": 10}).close()


def test_temporary_list_needs_the_shims_confirmation(lib):
    # sql.rs:70-86: a configured temporary that Resource does not hold is a construction error
    with pytest.raises(ArkError) as e:
        SqlProcessor({"query": "SELECT * FROM flow", "temporary_list": [{"name": "redis_temporary", "table_name": "redis_table", "key": {"type": "value", "value": "test"}}]})
    assert e.value.kind == "Process" and e.value.message == "Temporary redis_temporary not found"


def test_expr_config_and_evaluate_result_host_logic():
    # expr/mod.rs:30-49, 178-189: the serde-tagged Expr enum and EvaluateResult::get (no device involved)
    from arkflow_b200.expr import EvaluateResult, Expr

    e = Expr.from_config({"type": "expr", "expr": "concat(name, '!')"})
    assert e.kind == "Expr" and e.payload == "concat(name, '!')"
    v = Expr.from_config({"type": "value", "value": "test"})
    assert v.kind == "Value" and v.evaluate_expr(None).get(7) == "test"
    for bad in ({}, {"type": "expr"}, {"type": "nope", "value": 1}, "x"):
        with pytest.raises(ArkError) as err:
            Expr.from_config(bad)
        assert err.value.kind == "Serialization"
    s = EvaluateResult("Scalar", "test")
    assert s.get(0) == "test" and s.get(1) == "test"
    vec = EvaluateResult("Vec", ["a", "b"])
    assert vec.get(0) == "a" and vec.get(1) == "b" and vec.get(2) is None


@pytest.mark.parametrize("kind", [0, 1, 2, -1])
def test_host_copy_kinds_copy_exactly(lib, kind):
    """The staging copy of pageable inputs (csrc/host_copy.cpp): every kind the CPU supports must copy byte-exactly for any
    size, source misalignment (a sliced Arrow buffer) and destination misalignment, and must not touch bytes outside the
    destination range.  Pure host code: runs without a GPU."""
    import ctypes as C

    import numpy as np

    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, (6 << 20) + 512, dtype=np.uint8)
    sizes = [0, 1, 63, 64, 65, 255, 4095, 4096, 4097, 4096 + 255, 65536 + 3, (4 << 20) - 64, (4 << 20) + 129]
    used = set()
    for n in sizes:
        for s_off, d_off in ((0, 0), (8, 0), (3, 0), (0, 16), (5, 37), (64, 63)):
            dst = np.full(n + 256, 0xA5, dtype=np.uint8)
            got = lib.ark_host_copy(C.c_void_p(dst.ctypes.data + d_off), C.c_void_p(src.ctypes.data + s_off), n, kind)
            assert got >= 0
            used.add(got)
            assert np.array_equal(dst[d_off:d_off + n], src[s_off:s_off + n]), (kind, n, s_off, d_off)
            assert (dst[:d_off] == 0xA5).all() and (dst[d_off + n:] == 0xA5).all(), ("bytes outside the range were written", kind, n, s_off, d_off)
    if kind == 0:
        assert used == {0}
    assert lib.ark_host_copy(None, None, 8, kind) == -1


def test_input_config_errors(lib):
    """`generate` / `file` input builders (csrc/inputs.cu): config shape errors are serde-style `Error::Serialization`s as in the
    reference (`input/generate.rs:30-43`, `input/file.rs:395-455`), formats this library does not decode are Unsupported — all
    decided at construction, without a GPU."""
    from arkflow_b200.input import FileInput, GenerateInput, build_input

    GenerateInput({"context": "x", "interval": "1s", "batch_size": 3}).close()
    GenerateInput({"context": "{}", "interval": "10ms"}).close()  # batch_size and count are optional
    for cfg, frag in (({"interval": "1s"}, "missing field `context`"), ({"context": "x", "interval": "abc"}, "duration"),
                      ({"context": "x", "interval": "1ms", "count": -1}, "count"), ({"context": "x"}, "interval")):
        with pytest.raises(ArkError) as e:
            GenerateInput(cfg)
        assert e.value.kind == "Serialization" and frag in e.value.message, (cfg, e.value.message)
    for cfg, kind, frag in (({"path": "/tmp/x.csv"}, "Serialization", "input_type"), ({"input_type": {"type": "csv"}}, "Serialization", "path"),
                            ({"input_type": {"type": "xml", "path": "/tmp/x"}}, "Serialization", "unknown variant `xml`"),
                            ({"input_type": {"type": "parquet", "path": "/tmp/x.parquet"}}, "Unsupported", "parquet")):
        with pytest.raises(ArkError) as e:
            FileInput(cfg)
        assert e.value.kind == kind and frag in e.value.message, (cfg, e.value.message)
    FileInput({"input_type": {"type": "csv", "path": "/nonexistent.csv"}}).close()  # the file is opened by connect(), as in the reference
    with pytest.raises(ArkError):
        build_input({"type": "kafka"})
