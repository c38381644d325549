"""Builds tests/golden/reference_pins.json — every numeric / behavioural assertion the reference's own
tests hold for the hot path, restated on the SQL/type subset this library
accepts.  The reference cannot be executed here (Rust on un-vendored DataFusion 47; no cargo), so the
expected values are transcribed from the reference's test assertions, each citing file:line.  Where
the reference test uses Int32 / metadata columns / ORDER BY the pin is restated on Int64 columns and
the note says so.  Run:  python tests/golden/make_golden.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

PINS = [
    {"id": "sql_basic_query", "source": "crates/arkflow-plugin/src/processor/sql.rs:257-297",
     "columns": {"id": {"type": "int64", "values": [1, 2, 3]}, "name": {"type": "utf8", "values": ["a", "b", "c"]}},
     "query": "SELECT * FROM flow", "expect": {"kind": "Single", "rows": 3, "cols": 2}},
    {"id": "sql_empty_batch_is_none", "source": "crates/arkflow-plugin/src/processor/sql.rs:299-322",
     "columns": {"id": {"type": "int64", "values": []}, "name": {"type": "utf8", "values": []}},
     "query": "SELECT * FROM flow", "expect": {"kind": "None"}},
    {"id": "sql_invalid_query", "source": "crates/arkflow-plugin/src/processor/sql.rs:324-339",
     "columns": None, "query": "INVALID SQL QUERY", "expect": {"kind": "ConstructError", "error_kind": "Process", "prefix": "SQL query error"}},
    {"id": "sql_custom_table_name", "source": "crates/arkflow-plugin/src/processor/sql.rs:341-375",
     "columns": {"id": {"type": "int64", "values": [1]}}, "table_name": "custom_table",
     "query": "SELECT * FROM custom_table", "expect": {"kind": "Single", "rows": 1, "cols": 1}},
    {"id": "sql_pool_performance_query", "source": "crates/arkflow-plugin/src/processor/sql.rs:377-425",
     "columns": {"id": {"type": "int64", "values": [1, 2, 3, 4, 5]}, "value": {"type": "int64", "values": [10, 20, 30, 40, 50]}},
     "query": "SELECT * FROM flow WHERE id > 0", "repeat": 10, "expect": {"kind": "Single", "rows": 5, "cols": 2}},
    {"id": "count_star_is_int64_5", "source": "crates/arkflow-core/src/lib.rs:1811-1858",
     "note": "reference uses an Int32 id column + metadata columns on a raw SessionContext; COUNT(*) does not read them",
     "columns": {"id": {"type": "int64", "values": [1, 2, 3, 4, 5]}},
     "query": "SELECT COUNT(*) as total_count FROM flow",
     "expect": {"kind": "Single", "rows": 1, "cols": 1, "names": ["total_count"], "types": ["int64"], "values": {"total_count": [5]}}},
    {"id": "filter_cardinality_value_ge_150", "source": "crates/arkflow-core/src/lib.rs:2148-2197",
     "note": "reference wraps the filter in a subquery with ORDER BY over Int32 columns; restated as the inner filter on Int64",
     "columns": {"id": {"type": "int64", "values": [1, 2, 3]}, "value": {"type": "int64", "values": [100, 200, 300]}},
     "query": "SELECT id, value FROM flow WHERE value >= 150", "expect": {"kind": "Single", "rows": 2, "cols": 2, "values": {"value": [200, 300]}}},
    {"id": "missing_config", "source": "crates/arkflow-plugin/src/processor/sql.rs:235-239",
     "columns": None, "query": None, "expect": {"kind": "ConstructError", "error_kind": "Config", "prefix": "Batch processor configuration is missing"}},
    {"id": "generate_example_pipeline_sql", "source": "examples/generate_example.yaml:22-26 (derived, SURVEY.md §8(c): unverified by the reference)",
     "columns": {"timestamp": {"type": "int64", "values": [1625000000000]}, "value": {"type": "int64", "values": [10]}, "sensor": {"type": "utf8", "values": ["temp_1"]}},
     "query": "SELECT sum(value),avg(value) ,111 as x FROM flow  group by sensor",
     "expect": {"kind": "Single", "rows": 1, "cols": 3, "names": ["sum(flow.value)", "avg(flow.value)", "x"], "types": ["int64", "double", "int64"],
                "values": {"sum(flow.value)": [10], "avg(flow.value)": [10.0], "x": [111]}}},
    {"id": "stream_data_group_by", "source": "examples/stream_data.json:1-21 (sums derived in SURVEY.md §8(c), not asserted by the reference)",
     "fixture": "stream_data.json",
     "query": "SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor",
     "expect": {"kind": "Single", "rows": 2, "cols": 3, "unordered": True,
                "values": {"sensor": ["temp_1", "temp_2"], "sum(flow.value)": [223, 288], "count(*)": [11, 10]}}},
    {"id": "readme_quickstart_filter", "source": "README.md:58-79",
     "fixture": "stream_data.json", "query": "SELECT * FROM flow WHERE value >= 10",
     "expect": {"kind": "Single", "rows": 21, "cols": 3}},
]

if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(PINS, f, indent=1)
    print(f"wrote {len(PINS)} pins")
