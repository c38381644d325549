"""Host-side mirror of the reference's buffer stage for the hot path.

    arrow::compute::concat_batches   → concat_batches()      (memory.rs:130, window.rs:131,159)
    trait Buffer                     → class Buffer          (core/buffer/mod.rs:26-37)
    memory / session_window / tumbling_window builders → build_buffer(cfg)

All compute (concatenation, join) goes through the C ABI; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import itertools
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


class Ack:
    """trait Ack (core/input/mod.rs:52-57)."""

    def ack(self) -> None:  # pragma: no cover - interface
        pass


class NoopAck(Ack):
    pass


class VecAck(Ack):
    """core/input/mod.rs:66-95 / ArrayAck memory.rs:227-237: acknowledges every contained ack."""

    def __init__(self, acks):
        self.acks = list(acks)

    def ack(self) -> None:
        for a in self.acks:
            a.ack()


class Buffer:
    """trait Buffer (core/buffer/mod.rs:26-37) over the C ABI.  write(msg, ack) / read() / flush() / close()."""

    KIND = ""

    def __init__(self, config: Optional[dict], input_names: Optional[list] = None):
        lib = L.lib()
        handle = C.c_void_p()
        cfg = None if config is None else json.dumps(config).encode()
        names = None if input_names is None else json.dumps(list(input_names)).encode()
        _check(lib.ark_buffer_create(self.KIND.encode(), cfg, names, C.byref(handle)))
        self._h = handle
        self._acks: dict[int, Ack] = {}
        self._next = itertools.count(1)  # write() may be called from several input threads

    def write(self, msg, ack: Optional[Ack] = None) -> None:
        mb = msg if isinstance(msg, MessageBatch) else MessageBatch(msg)
        token = next(self._next)
        self._acks[token] = ack or NoopAck()
        arr, sch = F.export_record_batch(mb.record_batch)
        try:
            status = L.lib().ark_buffer_write(self._h, C.byref(arr), C.byref(sch),
                                              None if mb.input_name is None else mb.input_name.encode(), token)
        finally:
            F.release_schema(sch)
            F.release_array(arr)
        _check(status)

    def read(self):
        """Blocks like Buffer::read.  Returns None (closed and empty) or (MessageBatch, VecAck)."""
        out_arr, out_sch = L.ArrowArray(), L.ArrowSchema()
        cap = max(len(self._acks), 1) + 16
        acks = (C.c_uint64 * cap)()
        n = C.c_int64(0)
        _check(L.lib().ark_buffer_read(self._h, C.byref(out_arr), C.byref(out_sch), acks, cap, C.byref(n)))
        if not out_arr.release:
            return None
        rb = F.import_record_batch(out_arr, out_sch)
        got = [self._acks.pop(int(acks[i])) for i in range(n.value)]
        return MessageBatch(rb), VecAck(got)

    def write_device(self, batch: F.DeviceBatch, input_name: Optional[str] = None, ack: Optional[Ack] = None) -> None:
        """The same write for a batch that is already in HBM (ark_buffer_write_device: nothing is copied)."""
        token = next(self._next)
        self._acks[token] = ack or NoopAck()
        dev, sch = batch.export()
        try:
            status = L.lib().ark_buffer_write_device(self._h, C.byref(dev), C.byref(sch), None if input_name is None else input_name.encode(), token)
        finally:
            F.release_schema(sch)
            F.release_array(dev.array)
        _check(status)

    def read_device(self):
        """Like read(), the window stays in HBM: None or (DeviceBatch, VecAck)."""
        out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
        cap = max(len(self._acks), 1) + 16
        acks = (C.c_uint64 * cap)()
        n = C.c_int64(0)
        _check(L.lib().ark_buffer_read_device(self._h, C.byref(out_dev), C.byref(out_sch), acks, cap, C.byref(n)))
        if not out_dev.array.release:
            return None
        got = [self._acks.pop(int(acks[i]), NoopAck()) for i in range(n.value)]
        return F.DeviceBatch.adopt(out_dev, out_sch), VecAck(got)

    def flush(self) -> None:
        _check(L.lib().ark_buffer_flush(self._h))

    def close(self) -> None:
        _check(L.lib().ark_buffer_close(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().ark_buffer_destroy(self._h)
                self._h = None
        except Exception:
            pass


class MemoryBuffer(Buffer):
    """`type: memory` {capacity, timeout} — buffer/memory.rs:39-237."""

    KIND = "memory"


class SessionWindow(Buffer):
    """`type: session_window` {gap, join?} — buffer/session_window.rs:39-159."""

    KIND = "session_window"


class TumblingWindow(Buffer):
    """`type: tumbling_window` {interval, join?} — buffer/tumbling_window.rs:38-145."""

    KIND = "tumbling_window"


class SlidingWindow(Buffer):
    """`type: sliding_window` {window_size, interval, slide_size} — buffer/sliding_window.rs:37-238.
    A window is the first `window_size` queued BATCHES in arrival order; `slide_size` of them are then
    dropped, so consecutive windows overlap (and a batch's ack can be returned more than once)."""

    KIND = "sliding_window"

    def read(self):
        out_arr, out_sch = L.ArrowArray(), L.ArrowSchema()
        cap = max(len(self._acks), 1) + 16
        acks = (C.c_uint64 * cap)()
        n = C.c_int64(0)
        _check(L.lib().ark_buffer_read(self._h, C.byref(out_arr), C.byref(out_sch), acks, cap, C.byref(n)))
        if not out_arr.release:
            return None
        rb = F.import_record_batch(out_arr, out_sch)
        got = [self._acks[int(acks[i])] for i in range(n.value)]  # kept: the batch may be part of the next window too
        return MessageBatch(rb), VecAck(got)


_BUFFERS = {"memory": MemoryBuffer, "session_window": SessionWindow, "tumbling_window": TumblingWindow,
            "sliding_window": SlidingWindow}


def build_buffer(config: dict, input_names: Optional[list] = None) -> Buffer:
    """BufferConfig::build (core/buffer/mod.rs:50-73): {"type": ..., <flattened config>}."""
    cfg = dict(config)
    kind = cfg.pop("type", None)
    cfg.pop("name", None)
    if kind not in _BUFFERS:
        raise ArkError(L.ARK_ERR_CONFIG, f"Unknown buffer type: {kind}")
    return _BUFFERS[kind](cfg if cfg else None, input_names)
