"""Shared loader for tests/golden/reference_pins.json."""
import json
import os

import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
_T = {"int64": pa.int64(), "utf8": pa.utf8(), "float64": pa.float64(), "bool": pa.bool_(), "binary": pa.binary()}


def load_pins():
    return json.load(open(os.path.join(HERE, "golden", "reference_pins.json")))


def stream_data_batch():
    rows = [json.loads(l) for l in open(os.path.join(HERE, "golden", "stream_data.json"))]
    return pa.record_batch({"timestamp": pa.array([r["timestamp"] for r in rows], pa.int64()),
                            "value": pa.array([r["value"] for r in rows], pa.int64()),
                            "sensor": pa.array([r["sensor"] for r in rows])})


def pin_batch(pin):
    if pin.get("fixture") == "stream_data.json":
        return stream_data_batch()
    if pin.get("columns") is None:
        return None
    return pa.record_batch({k: pa.array(v["values"], _T[v["type"]]) for k, v in pin["columns"].items()})


def check_expect(pin, result):
    """result: None (ProcessResult::None) or a pyarrow RecordBatch."""
    exp = pin["expect"]
    if exp["kind"] == "None":
        assert result is None, pin["id"]
        return
    assert result is not None, pin["id"]
    assert result.num_rows == exp["rows"], (pin["id"], result.num_rows)
    assert result.num_columns == exp["cols"], (pin["id"], result.num_columns)
    if "names" in exp:
        assert result.schema.names == exp["names"], (pin["id"], result.schema.names)
    if "types" in exp:
        assert [str(f.type) for f in result.schema] == exp["types"], (pin["id"], result.schema)
    if "values" in exp:
        got = {n: result.column(n).to_pylist() for n in exp["values"]}
        if exp.get("unordered"):
            names = list(exp["values"])
            assert sorted(zip(*[got[n] for n in names])) == sorted(zip(*[exp["values"][n] for n in names])), pin["id"]
        else:
            assert got == exp["values"], (pin["id"], got)
