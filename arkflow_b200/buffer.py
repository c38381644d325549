"""Host-side mirror of the reference's buffer stage for the hot path.

    arrow::compute::concat_batches   → concat_batches()      (memory.rs:130, window.rs:131,159)
    trait Buffer                     → class Buffer          (core/buffer/mod.rs:26-37)
    memory / session_window / tumbling_window builders → build_buffer(cfg)

All compute (concatenation, join) goes through the C ABI; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Optional

import pyarrow as pa

from . import _lib as L
from . import arrow_ffi as F
from .processor import ArkError, MessageBatch, _check


def concat_batches(batches: list[pa.RecordBatch]) -> pa.RecordBatch:
    """Device concat of host RecordBatches (H2D, segmented-copy kernels, D2H)."""
    lib = L.lib()
    n = len(batches)
    arrs = (L.ArrowArray * max(n, 1))()
    schs = (L.ArrowSchema * max(n, 1))()
    for i, rb in enumerate(batches):
        rb._export_to_c(C.addressof(arrs[i]), C.addressof(schs[i]))
    out_arr, out_sch = L.ArrowArray(), L.ArrowSchema()
    try:
        status = lib.ark_concat_batches(n, arrs, schs, C.byref(out_arr), C.byref(out_sch))
    finally:
        for i in range(n):
            F.release_schema(schs[i])
            F.release_array(arrs[i])
    _check(status)
    return F.import_record_batch(out_arr, out_sch)


def concat_batches_device(batches: list[F.DeviceBatch]) -> F.DeviceBatch:
    lib = L.lib()
    n = len(batches)
    devs = (L.ArrowDeviceArray * max(n, 1))()
    schs = (L.ArrowSchema * max(n, 1))()
    for i, b in enumerate(batches):
        d, s = b.export()
        C.memmove(C.addressof(devs[i]), C.addressof(d), C.sizeof(L.ArrowDeviceArray))
        C.memmove(C.addressof(schs[i]), C.addressof(s), C.sizeof(L.ArrowSchema))
    out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
    try:
        status = lib.ark_concat_batches_device(n, devs, schs, C.byref(out_dev), C.byref(out_sch))
    finally:
        for i in range(n):
            F.release_schema(schs[i])
            F.release_array(devs[i].array)
    _check(status)
    return F.DeviceBatch.adopt(out_dev, out_sch)
