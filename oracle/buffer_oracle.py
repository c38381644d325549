"""CPU oracle for the buffer stage's data path — TEST INFRASTRUCTURE (see oracle/__init__.py).

Only the *data* semantics are restated (what a drained window contains and in which row order); the
trigger logic (timers, capacity) is host code tested directly against the reference's own tests.
  memory buffer      : drain oldest-first, concat            buffer/memory.rs:106-138,155
  window buffers     : per input drain newest-first, concat per input, then concat across inputs in
                       first-seen input order, or JoinOperation                 buffer/window.rs:99-190
  JoinOperation      : json-decode each input's __value__, require every configured input, run the SQL
                                                                                 buffer/join.rs:62-146
"""
from __future__ import annotations

from typing import Optional

import pyarrow as pa

from .json_oracle import json_to_arrow
from .sql_oracle import sql_join


def concat(batches: list[pa.RecordBatch]) -> pa.RecordBatch:
    t = pa.Table.from_batches(batches).combine_chunks()
    b = t.to_batches()
    return b[0] if b else pa.RecordBatch.from_arrays([pa.array([], f.type) for f in t.schema], schema=t.schema)


def memory_drain(writes: list[pa.RecordBatch]) -> pa.RecordBatch:
    return concat(list(writes))  # FIFO


def window_drain(writes: list[tuple[Optional[str], pa.RecordBatch]], join: Optional[dict] = None, input_names=()) -> pa.RecordBatch:
    order, queues = [], {}
    for name, rb in writes:
        name = name or ""
        if name not in queues:
            queues[name] = []
            order.append(name)
        queues[name].insert(0, rb)  # push_front
    per_input = [(n, concat(queues[n])) for n in order]  # pop_front ⇒ newest first
    if join is None:
        return concat([b for _, b in per_input])
    tables = {}
    for n, b in per_input:
        decoded = json_to_arrow(b, join.get("value_field") or "__value__")
        if n:
            tables[n] = decoded
    if not all(n in tables for n in input_names):
        return pa.RecordBatch.from_arrays([], schema=pa.schema([]))
    return sql_join(tables, join["query"])


def sliding_windows(writes: list[pa.RecordBatch], window_size: int, slide_size: int) -> list[pa.RecordBatch]:
    """SlidingWindow::process_slide repeated until fewer than window_size batches remain
    (buffer/sliding_window.rs:104-158): a window is the first window_size queued batches concatenated in
    arrival order; then slide_size batches leave the front of the queue.  PARITY: the reference asserts
    only that a window is produced (sliding_window.rs:396-418) — contents unpinned."""
    q = list(writes)
    out = []
    while len(q) >= window_size:
        out.append(concat(q[:window_size]))
        del q[:slide_size]
    return out


def batch_processor(writes: list[pa.RecordBatch], count: int) -> list[pa.RecordBatch]:
    """BatchProcessor::process by count only (processor/batch.rs:95-116): every `count`-th message returns
    the concatenation of the held ones.  batch.rs:170-186 pins `batch.len() == 2` for count = 2."""
    out, held = [], []
    for rb in writes:
        held.append(rb)
        if len(held) >= count:
            out.append(concat(held))
            held = []
    return out
