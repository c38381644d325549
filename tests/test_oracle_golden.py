"""The oracle against every pin the reference's own tests hold for the sql processor
(tests/golden/reference_pins.json ← tests/golden/make_golden.py).  CPU only."""
import pytest

from golden_util import check_expect, load_pins, pin_batch
from oracle.sql_oracle import OracleError, parse, sql_process

PINS = load_pins()


@pytest.mark.parametrize("pin", PINS, ids=[p["id"] for p in PINS])
def test_oracle_matches_reference_pin(pin):
    exp = pin["expect"]
    if exp["kind"] == "ConstructError":
        if pin["query"] is None:
            pytest.skip("configuration-missing is a builder concern (tested against the C ABI)")
        with pytest.raises(OracleError) as e:
            parse(pin["query"])
        assert e.value.kind == exp["error_kind"] and e.value.message.startswith(exp["prefix"])
        return
    rb = pin_batch(pin)
    for _ in range(pin.get("repeat", 1)):
        out = sql_process(rb, pin["query"], pin.get("table_name", "flow"))
        check_expect(pin, out)


def test_oracle_float_total_order():
    import pyarrow as pa

    rb = pa.record_batch({"v": pa.array([float("nan"), float("inf"), -0.0, 0.0, 10.0, 9.5], pa.float64()), "i": pa.array(range(6), pa.int64())})
    assert sql_process(rb, "SELECT i FROM flow WHERE v >= 10").column(0).to_pylist() == [0, 1, 4]   # NaN, +inf, 10.0
    assert sql_process(rb, "SELECT i FROM flow WHERE v < 0").column(0).to_pylist() == [2]             # -0.0 < +0.0
    assert sql_process(rb, "SELECT i FROM flow WHERE v = 0").column(0).to_pylist() == [3]             # bitwise equality


def test_oracle_wrapping_and_division():
    import pyarrow as pa

    rb = pa.record_batch({"v": pa.array([2**62, -(2**62), 7, -7], pa.int64())})
    out = sql_process(rb, "SELECT v * 4, v / 2, v % 3 FROM flow")
    assert out.column(0).to_pylist() == [0, 0, 28, -28]
    assert out.column(1).to_pylist() == [2**61, -(2**61), 3, -3]
    assert out.column(2).to_pylist() == [(2**62) % 3, -((2**62) % 3), 1, -1]
    with pytest.raises(OracleError):
        sql_process(rb, "SELECT v / 0 FROM flow")
