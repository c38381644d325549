"""concat_batches on the device vs arrow's own concat (the reference calls arrow::compute::concat_batches,
buffer/memory.rs:130, buffer/window.rs:131,159).  Byte-exact."""
import numpy as np
import pyarrow as pa
import pytest

from arkflow_b200.arrow_ffi import DeviceBatch
from arkflow_b200.buffer import concat_batches, concat_batches_device
from arkflow_b200.processor import ArkError
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def check(batches):
    want = pa.Table.from_batches(batches).combine_chunks()
    for got in (concat_batches(batches), concat_batches_device([DeviceBatch.from_arrow(b) for b in batches]).to_arrow()):
        assert got.schema.names == want.schema.names
        assert got.num_rows == want.num_rows
        for i, name in enumerate(want.schema.names):
            w = want.column(i).combine_chunks() if want.num_rows else pa.array([], want.schema.field(i).type)
            assert got.column(i).to_pylist() == w.to_pylist(), name
    return got


def test_concat_schema_s_batches(gpu):
    sizes = [1, 8191, 8192, 100_003, 7, 65_536]
    batches, row0 = [], 0
    for n in sizes:
        batches.append(synth_batch(n, row0=row0, key_space=999))
        row0 += n
    check(batches)


def test_concat_single_and_empty_members(gpu):
    a = synth_batch(1000)
    check([a])
    check([a.slice(0, 0), a, a.slice(10, 0), a.slice(5, 17)])


def test_concat_nulls_bools_ragged_strings_and_slices(gpu):
    rng = np.random.default_rng(1)
    def mk(n, seed, with_nulls):
        r = np.random.default_rng(seed)
        s = [None if (with_nulls and r.random() < 0.2) else "s" * int(r.integers(0, 30)) + str(i) for i in range(n)]
        return pa.record_batch({
            "i": pa.array(r.integers(-100, 100, n), pa.int64(), mask=(r.random(n) < 0.3) if with_nulls else None),
            "f": pa.array(r.random(n), pa.float64()),
            "b": pa.array([None if (with_nulls and r.random() < 0.1) else bool(x) for x in r.integers(0, 2, n)], pa.bool_()),
            "s": pa.array(s, pa.utf8()),
            "z": pa.array([(x or "").encode() for x in s], pa.binary()),
        })
    batches = [mk(1003, 1, True), mk(5, 2, False), mk(4099, 3, True).slice(13, 3001), mk(64, 4, False).slice(3, 50), mk(1, 5, True)]
    check(batches)


def test_concat_binary_payload_batches_config5(gpu):
    # config 5: raw `__value__` Binary payloads (generate input, 63-byte JSON messages)
    msg = b'{ "timestamp": 1625000000000, "value": 10, "sensor": "temp_1" }'
    batches = [pa.record_batch([pa.array([msg] * n, pa.binary())], schema=pa.schema([pa.field("__value__", pa.binary(), nullable=False)]))
               for n in (1000, 1, 4096, 333)]
    out = check(batches)
    assert out.num_rows == 5430


def test_concat_schema_mismatch_is_process_error(gpu):
    a = pa.record_batch({"x": pa.array([1], pa.int64())})
    b = pa.record_batch({"x": pa.array(["a"])})
    with pytest.raises(ArkError) as e:
        concat_batches([a, b])
    assert e.value.kind == "Process" and e.value.message.startswith("Merge batches failed")
