"""`sliding_window` buffer and `batch` processor over the C ABI.  The cases mirror the
reference's own tests (buffer/sliding_window.rs:290-467, processor/batch.rs:150-258); window contents are
checked against oracle/buffer_oracle.py (the reference asserts only Some(_) / row counts)."""
import threading
import time

import pyarrow as pa
import pytest

from arkflow_b200.buffer import Ack, SlidingWindow, build_buffer
from arkflow_b200.processor import ArkError, BatchProcessor, MessageBatch, build_processor, init
from oracle.buffer_oracle import batch_processor, sliding_windows
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def read_with_timeout(buf, timeout_s):
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("r", buf.read()), daemon=True)
    t.start()
    t.join(timeout_s)
    return (not t.is_alive()), box.get("r")


class CountAck(Ack):
    def __init__(self):
        self.n = 0

    def ack(self):
        self.n += 1


def msg(i):
    return MessageBatch.new_binary([f"msg{i}".encode()])


# ---- sliding_window ----------------------------------------------------------------------------------
def test_sliding_window_basic(gpu):
    # sliding_window.rs:396-418: window 3, slide 2; three writes → read yields a window
    buf = SlidingWindow({"window_size": 3, "interval": "100ms", "slide_size": 2})
    for i in range(3):
        buf.write(msg(i))
    done, r = read_with_timeout(buf, 1.0)
    assert done and r is not None
    assert r[0].record_batch.column("__value__").to_pylist() == [b"msg0", b"msg1", b"msg2"]


def test_sliding_window_write_and_close(gpu):
    # sliding_window.rs:420-444: two writes (< window_size), close → read returns Ok(None)
    buf = SlidingWindow({"window_size": 3, "interval": "100ms", "slide_size": 1})
    for i in range(2):
        buf.write(msg(i))
    buf.close()
    done, r = read_with_timeout(buf, 1.0)
    assert done and r is None


def test_sliding_window_flush(gpu):
    # sliding_window.rs:446-467
    buf = SlidingWindow({"window_size": 3, "interval": "100ms", "slide_size": 1})
    for i in range(2):
        buf.write(msg(i))
    buf.flush()
    done, r = read_with_timeout(buf, 1.0)
    assert done and r is None


def test_sliding_window_blocks_until_enough_batches(gpu):
    buf = SlidingWindow({"window_size": 3, "interval": "50ms", "slide_size": 1})
    buf.write(msg(0))
    buf.write(msg(1))
    done, _ = read_with_timeout(buf, 0.25)
    assert not done  # two batches never make a window, however many timer ticks pass
    buf.write(msg(2))
    time.sleep(0.3)
    buf.close()


def test_sliding_window_contents_overlap_and_acks(gpu):
    writes = [synth_batch(1000 + 17 * i, row0=5000 * i, key_space=50) for i in range(7)]
    acks = [CountAck() for _ in writes]
    buf = SlidingWindow({"window_size": 3, "interval": "20ms", "slide_size": 2})
    for rb, a in zip(writes, acks):
        buf.write(MessageBatch.new_arrow(rb), a)
    want = sliding_windows(writes, 3, 2)
    assert len(want) == 3  # [0,1,2] [2,3,4] [4,5,6]
    for w in want:
        done, r = read_with_timeout(buf, 2.0)
        assert done and r is not None
        assert r[0].record_batch.equals(w)
        r[1].ack()
    assert [a.n for a in acks] == [1, 1, 2, 1, 2, 1, 1]  # batches 2 and 4 sit in two windows each
    done, _ = read_with_timeout(buf, 0.2)
    assert not done  # one batch left: no window
    buf.close()


def test_sliding_window_builder_validation(gpu):
    # sliding_window.rs:318-394
    ok = build_buffer({"type": "sliding_window", "window_size": 10, "interval": "1s", "slide_size": 5})
    ok.close()
    for cfg in ({"window_size": 0, "interval": "1s", "slide_size": 5},
                {"window_size": 10, "interval": "1s", "slide_size": 0},
                {"window_size": 5, "interval": "1s", "slide_size": 10}):
        with pytest.raises(ArkError) as e:
            build_buffer({"type": "sliding_window", **cfg})
        assert e.value.kind == "Config"
    with pytest.raises(ArkError) as e:
        SlidingWindow(None)
    assert e.value.kind == "Config" and "Sliding window configuration is missing" in str(e.value)


# ---- batch processor -----------------------------------------------------------------------------------
def test_batch_processor_size(gpu):
    # batch.rs:153-186
    p = BatchProcessor({"count": 2, "timeout_ms": 1000})
    assert p.process(MessageBatch.new_binary([b"test1"])).is_empty()
    r = p.process(MessageBatch.new_binary([b"test2"]))
    assert r.kind == "Single" and r.batches[0].num_rows == 2
    assert r.batches[0].record_batch.column("__value__").to_pylist() == [b"test1", b"test2"]


def test_batch_processor_timeout(gpu):
    # batch.rs:188-224
    p = BatchProcessor({"count": 5, "timeout_ms": 100})
    assert p.process(MessageBatch.new_binary([b"test1"])).is_empty()
    time.sleep(0.15)
    r = p.process(MessageBatch.new_binary([b"test2"]))
    assert r.kind == "Single" and r.batches[0].num_rows == 2


def test_batch_processor_empty_flush_and_close(gpu):
    # batch.rs:226-258
    p = BatchProcessor({"count": 2, "timeout_ms": 1000})
    assert p.flush().is_empty()
    q = BatchProcessor({"count": 5, "timeout_ms": 1000})
    q.process(MessageBatch.new_binary([b"test1"]))
    q.close()
    assert q.flush().is_empty()


def test_batch_processor_contents_and_registry(gpu):
    init()
    p = build_processor({"type": "batch", "count": 3, "timeout_ms": 60000})
    writes = [synth_batch(500 + i, row0=1000 * i, key_space=9, value_kind=0) for i in range(7)]
    got = []
    for rb in writes:
        r = p.process(MessageBatch.new_arrow(rb))
        got.extend(b.record_batch for b in r.batches)
    want = batch_processor(writes, 3)
    assert len(got) == len(want) == 2
    for g, w in zip(got, want):
        assert g.equals(w)
    tail = p.flush()
    assert tail.batches[0].record_batch.equals(writes[6])
    with pytest.raises(ArkError) as e:
        BatchProcessor(None)
    assert e.value.kind == "Config" and "Batch processor configuration is missing" in str(e.value)
    with pytest.raises(ArkError) as e:
        BatchProcessor({"count": 2})
    assert e.value.kind == "Serialization"
