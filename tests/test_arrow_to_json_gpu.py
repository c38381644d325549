"""`arrow_to_json` on the device vs the oracle (mirrors crates/arkflow-plugin/src/processor/json.rs:268-343),
plus the whole shipped example pipeline of examples/generate_example.yaml."""
import json
import struct

import numpy as np
import pyarrow as pa
import pytest

from arkflow_b200.processor import ArkError, ArrowToJsonProcessor, JsonToArrowProcessor, MessageBatch, Pipeline, SqlProcessor
from oracle.json_oracle import arrow_to_json, arrow_to_json_lines, json_to_arrow
from oracle.sql_oracle import sql_process as sql_oracle_process

pytestmark = pytest.mark.gpu


def run(rb, cfg=None):
    return ArrowToJsonProcessor(cfg or {}).process(MessageBatch.new_arrow(rb)).batches[0].record_batch


def check(rb, cfg=None):
    inc = set(cfg["fields_to_include"]) if cfg and "fields_to_include" in cfg else None
    want = arrow_to_json(rb, inc)
    got = run(rb, cfg)
    assert got.schema.names == want.schema.names
    assert got.num_rows == want.num_rows
    g, w = got.column(got.num_columns - 1).to_pylist(), want.column(want.num_columns - 1).to_pylist()
    for i, (a, b) in enumerate(zip(g, w)):
        assert a == b, (i, a, b)
    for i in range(rb.num_columns):
        a, b = got.column(i), rb.column(i)
        if b.type == pa.float64() and b.null_count == 0:  # Array.equals treats NaN != NaN: compare the bits
            assert a.null_count == 0
            assert np.array_equal(a.to_numpy(zero_copy_only=False).view(np.uint64), b.to_numpy(zero_copy_only=False).view(np.uint64))
        else:
            assert a.equals(b)
    return got


def test_arrow_to_json_basic(gpu):
    # json.rs:268-312: round trip of one decoded record
    rec = {"int_field": 42, "float_field": 3.14, "string_field": "hello", "bool_field": True}
    rb = json_to_arrow(MessageBatch.new_binary([json.dumps(rec).encode()]).record_batch)
    out = check(rb)
    assert out.num_rows == 1 and json.loads(out.column("__value__")[0].as_py()) == rec


def test_missing_config(gpu):
    # json.rs:314-343
    with pytest.raises(ArkError) as e:
        ArrowToJsonProcessor(None)
    assert e.value.kind == "Config"


def test_types_nulls_escapes_and_field_filter(gpu):
    rng = np.random.default_rng(3)
    n = 3000
    strs = ["plain", "", 'q"uote', "back\\slash", "tab\tnl\n", "ctl\x01\x1f", "unicode é 漢 😀", "/slash"]
    rb = pa.record_batch({
        "i": pa.array([None if rng.random() < 0.2 else int(x) for x in rng.integers(-2**62, 2**62, n)], pa.int64()),
        "f": pa.array(rng.normal(0, 1e3, n), pa.float64(), mask=rng.random(n) < 0.2),
        "s": pa.array([None if rng.random() < 0.2 else strs[int(k)] for k in rng.integers(0, len(strs), n)]),
        "b": pa.array([None if rng.random() < 0.2 else bool(k) for k in rng.integers(0, 2, n)], pa.bool_()),
        "z": pa.array([bytes(rng.integers(0, 256, int(k)).tolist()) for k in rng.integers(0, 6, n)], pa.binary()),
        'we"ird\tname': pa.array(range(n), pa.int64()),
    })
    check(rb)
    check(rb, {"fields_to_include": ["s", "i"]})
    allnull = pa.record_batch({"a": pa.array([None, None], pa.int64()), "n": pa.nulls(2)})
    assert arrow_to_json_lines(allnull) == [b"{}", b"{}"]
    check(allnull)


def test_float_formatting_matches_shortest_round_trip(gpu):
    rng = np.random.default_rng(11)
    special = [0.0, -0.0, 1.0, 10.0, 0.1, 0.3, 1e21, 1e22, 1e23, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, 123456789.125,
               20.272727272727273, 28.8, 4.35, 1e-5, 1e-6, 1234567890.0, 12345678901.0, float("nan"), float("inf"), -float("inf"), 9007199254740993.0]
    bits = rng.integers(0, 2**63, 20000, dtype=np.uint64) | (rng.integers(0, 2, 20000, dtype=np.uint64) << np.uint64(63))
    rnd = [struct.unpack("<d", struct.pack("<Q", int(b)))[0] for b in bits]
    scaled = (rng.random(20000) * 10.0 ** rng.integers(-12, 13, 20000)).tolist()
    vals = special + rnd + scaled
    rb = pa.record_batch({"f": pa.array(vals, pa.float64())})
    out = check(rb)
    for v, line in zip(vals, out.column("__value__").to_pylist()):
        if v == v and abs(v) != float("inf"):
            assert float(json.loads(line)["f"]) == v  # round-trips exactly


def test_generate_example_pipeline_end_to_end(gpu):
    # examples/generate_example.yaml:22-32 — json_to_arrow → sql (GROUP BY) → arrow_to_json → sql (cast to string)
    payload = b'{ "timestamp": 1625000000000, "value": 10, "sensor": "temp_1" }'
    pipe = Pipeline([
        JsonToArrowProcessor({}),
        SqlProcessor({"query": "SELECT sum(value),avg(value) ,111 as x FROM flow  group by sensor"}),
        ArrowToJsonProcessor({}),
        SqlProcessor({"query": "SELECT *,cast( __value__  as string) as y FROM flow "}),
    ])
    out = pipe.process(MessageBatch.new_binary([payload])).batches[0].record_batch
    assert out.schema.names == ["sum(flow.value)", "avg(flow.value)", "x", "__value__", "y"]
    # the stdout output prints each __value__ payload (output/stdout.rs:68-86): SURVEY.md §8(c)'s derived line
    assert out.column("__value__").to_pylist() == [b'{"sum(flow.value)":10,"avg(flow.value)":10.0,"x":111}']
    assert out.column("y").to_pylist() == ['{"sum(flow.value)":10,"avg(flow.value)":10.0,"x":111}']


def test_large_batch(gpu):
    from oracle.synth import synth_batch

    rb = synth_batch(100_000, value_kind=1, key_space=1000)
    check(rb)


def test_device_resident_chain_json_sql_json(gpu):
    # json_to_arrow → sql → arrow_to_json without leaving HBM; the batch crosses PCIe once each way
    from arkflow_b200.arrow_ffi import DeviceBatch

    payloads = [json.dumps({"timestamp": 1625000000000 + i, "value": i % 20, "sensor": f"temp_{i % 7}"}).encode() for i in range(5000)]
    host = MessageBatch.new_binary(payloads)
    dev = DeviceBatch.from_arrow(host.record_batch)
    q = "SELECT sensor, value FROM flow WHERE value >= 10"
    stages = [JsonToArrowProcessor({}), SqlProcessor({"query": q}), ArrowToJsonProcessor({})]
    cur = dev
    for st in stages:
        cur = st.process_device(cur)
    got = cur.to_arrow()
    want = arrow_to_json(sql_oracle_process(json_to_arrow(host.record_batch), q))
    assert got.schema.names == want.schema.names == ["sensor", "value", "__value__"]
    assert got.column("__value__").to_pylist() == want.column("__value__").to_pylist()
    assert got.num_rows == 2500
