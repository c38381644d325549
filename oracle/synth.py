"""Synthetic input of schema S — numpy twin of ark_synth_batch_device.

timestamp[i] = 1625000000000 + 1000·i                       (Int64; step of examples/stream_data.json)
value[i]     = r_v(i) mod 20                                  (Int64 uniform [0,20) ⇒ σ(value >= 10) = 0.5)
             | 20 · (r_v(i) >> 11) · 2⁻⁵³                     (Float64 variant)
sensor[i]    = "temp_%07d" % (r_k(i) mod K)                   (Utf8, 12 bytes)
r_v(i) = i-th output of splitmix64 seeded with `seed`, r_k(i) = same with seed ^ 0x9E37.
"""
from __future__ import annotations

import numpy as np
import pyarrow as pa

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)


def splitmix64_at(seed: int, idx: np.ndarray) -> np.ndarray:
    """Output number idx (0-based) of the splitmix64 stream seeded with `seed` (uint64 arithmetic)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        return z ^ (z >> np.uint64(31))


def synth_columns(n_rows: int, row0: int = 0, seed: int = 42, value_kind: int = 0, key_space: int = 1_000_000):
    idx = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    ts = (np.int64(1625000000000) + np.int64(1000) * idx.astype(np.int64)).astype(np.int64)
    rv = splitmix64_at(seed, idx)
    if value_kind == 0:
        value = (rv % np.uint64(20)).astype(np.int64)
    else:
        value = (rv >> np.uint64(11)).astype(np.float64) * (20.0 * 2.0 ** -53)
    key = (splitmix64_at(seed ^ 0x9E37, idx) % np.uint64(key_space)).astype(np.int64)
    return ts, value, key


def sensor_array(key: np.ndarray) -> pa.Array:
    """"temp_%07d" % key as a Utf8 array built from raw buffers (fast for millions of rows)."""
    n = len(key)
    digits = np.empty((n, 12), dtype=np.uint8)
    digits[:, :5] = np.frombuffer(b"temp_", dtype=np.uint8)
    k = key.astype(np.int64).copy()
    for d in range(6, -1, -1):
        digits[:, 5 + d] = (k % 10).astype(np.uint8) + ord("0")
        k //= 10
    offsets = (np.arange(n + 1, dtype=np.int64) * 12).astype(np.int32)
    return pa.Array.from_buffers(pa.utf8(), n, [None, pa.py_buffer(offsets.tobytes()), pa.py_buffer(digits.tobytes())])


def synth_batch(n_rows: int, row0: int = 0, seed: int = 42, value_kind: int = 0, key_space: int = 1_000_000) -> pa.RecordBatch:
    ts, value, key = synth_columns(n_rows, row0, seed, value_kind, key_space)
    return pa.RecordBatch.from_arrays(
        [pa.array(ts, type=pa.int64()), pa.array(value), sensor_array(key)],
        schema=pa.schema([pa.field("timestamp", pa.int64()), pa.field("value", pa.int64() if value_kind == 0 else pa.float64()),
                          pa.field("sensor", pa.utf8())]))
