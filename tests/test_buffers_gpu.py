"""Buffer stage over the C ABI: wake-up rules mirror the reference's own tests
(buffer/memory.rs:279-372, session_window.rs:209-416, tumbling_window.rs:196-350); window contents and
row order are checked against oracle/buffer_oracle.py (the reference asserts only Some(_)/is_ok())."""
import threading
import time

import pyarrow as pa
import pytest

from arkflow_b200.buffer import Ack, MemoryBuffer, SessionWindow, TumblingWindow, build_buffer
from arkflow_b200.processor import ArkError, MessageBatch
from oracle.buffer_oracle import memory_drain, window_drain
from oracle.synth import synth_batch

pytestmark = pytest.mark.gpu


def read_with_timeout(buf, timeout_s):
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("r", buf.read()), daemon=True)
    t0 = time.perf_counter()
    t.start()
    t.join(timeout_s)
    return (not t.is_alive()), box.get("r"), time.perf_counter() - t0


class CountAck(Ack):
    def __init__(self):
        self.n = 0

    def ack(self):
        self.n += 1


def bin_batch(*payloads):
    return MessageBatch.new_binary(list(payloads))


def test_memory_buffer_capacity_limit(gpu):
    # memory.rs:280-297: capacity 2, timeout 100ms; two 1-row writes → read returns within 200ms
    buf = MemoryBuffer({"capacity": 2, "timeout": "100ms"})
    buf.write(bin_batch(b"a"))
    buf.write(bin_batch(b"b"))
    done, r, _ = read_with_timeout(buf, 0.2)
    assert done and r is not None and r[0].num_rows == 2


def test_memory_buffer_timeout_notify(gpu):
    # memory.rs:299-312
    buf = MemoryBuffer({"capacity": 10, "timeout": "100ms"})
    buf.write(bin_batch(b"a"))
    done, r, _ = read_with_timeout(buf, 0.3)
    assert done and r is not None


def test_memory_buffer_waits_for_timer_when_empty(gpu):
    buf = MemoryBuffer({"capacity": 10, "timeout": "150ms"})
    threading.Timer(0.02, lambda: buf.write(bin_batch(b"late"))).start()
    done, r, dt = read_with_timeout(buf, 1.0)
    assert done and r is not None and dt >= 0.1  # woken by the timer tick, not by the sub-capacity write


def test_memory_buffer_flush_and_close(gpu):
    # memory.rs:314-344
    buf = MemoryBuffer({"capacity": 10, "timeout": "10s"})
    buf.write(bin_batch(b"a"))
    buf.flush()
    done, r, _ = read_with_timeout(buf, 0.2)
    assert done and r is not None
    buf2 = MemoryBuffer({"capacity": 10, "timeout": "10s"})
    buf2.close()
    done, r, _ = read_with_timeout(buf2, 0.2)
    assert done and r is None


def test_memory_buffer_concurrent_write_read_and_acks(gpu):
    # memory.rs:346-372: 10 single-row writes from another thread; reads sum to 10 rows
    buf = MemoryBuffer({"capacity": 100, "timeout": "100ms"})
    acks = [CountAck() for _ in range(10)]

    def writer():
        for i in range(10):
            buf.write(bin_batch(b"m%d" % i), acks[i])
            time.sleep(0.005)

    threading.Thread(target=writer, daemon=True).start()
    total, t0 = 0, time.time()
    while total < 10 and time.time() - t0 < 5:
        done, r, _ = read_with_timeout(buf, 0.5)
        if done and r is not None:
            total += r[0].num_rows
            r[1].ack()
    assert total == 10 and all(a.n == 1 for a in acks)


def test_memory_buffer_contents_fifo(gpu):
    writes = [synth_batch(n, row0=i * 1000, key_space=50) for i, n in enumerate([10, 1, 300, 7])]
    buf = MemoryBuffer({"capacity": 1000, "timeout": "10s"})
    for w in writes:
        buf.write(MessageBatch.new_arrow(w))
    buf.flush()
    got = buf.read()[0].record_batch
    assert got.equals(memory_drain(writes))


def test_window_contents_lifo_per_input(gpu):
    a = [synth_batch(n, row0=i * 100, key_space=5) for i, n in enumerate([3, 5, 2])]
    b = [synth_batch(n, row0=1000 + i * 100, key_space=5) for i, n in enumerate([4, 1])]
    writes = [("in1", a[0]), ("in2", b[0]), ("in1", a[1]), ("in2", b[1]), ("in1", a[2])]
    win = TumblingWindow({"interval": "50ms"})
    for name, rb in writes:
        win.write(MessageBatch(rb, name))
    done, r, _ = read_with_timeout(win, 1.0)
    assert done and r is not None
    assert r[0].record_batch.equals(window_drain(writes))


def test_tumbling_window_emits_on_tick_and_closes(gpu):
    # tumbling_window.rs:196-318
    win = TumblingWindow({"interval": "100ms"})
    threading.Timer(0.01, lambda: win.write(bin_batch(b"x"))).start()
    done, r, dt = read_with_timeout(win, 1.0)
    assert done and r is not None and dt >= 0.08
    win.close()
    done, r, _ = read_with_timeout(win, 0.2)
    assert done and r is None


def test_session_window_waits_for_gap(gpu):
    # session_window.rs:209-329: emits only once `gap` has elapsed since the last write
    win = SessionWindow({"gap": "150ms"})
    win.write(bin_batch(b"a"))
    t = threading.Timer(0.08, lambda: win.write(bin_batch(b"b")))
    t.start()
    done, r, dt = read_with_timeout(win, 2.0)
    assert done and r is not None and r[0].num_rows == 2
    assert dt >= 0.2  # 80 ms until the second write + the 150 ms gap
    win.flush()
    done, r, _ = read_with_timeout(win, 0.2)
    assert done and r is None


def test_window_flush_drains_waiting_reader(gpu):
    win = SessionWindow({"gap": "10s"})
    win.write(bin_batch(b"a"))
    threading.Timer(0.05, win.flush).start()
    done, r, _ = read_with_timeout(win, 1.0)
    assert done and r is not None and r[0].num_rows == 1


def test_join_window_example(gpu):
    # examples/join_buffer_example.yaml: two generate inputs, session window with a json-codec join
    m1 = b'{ "id": 1625000000000, "value": 10, "sensor": "temp_1" }'
    m2 = b'{ "id": 1625000000000, "value": 20, "sensor": "temp_2" }'
    join = {"query": "SELECT * FROM flow_input1 join flow_input2 on (flow_input1.id = flow_input2.id)", "codec": {"type": "json"}}
    names = ["flow_input1", "flow_input2"]
    win = build_buffer({"type": "session_window", "gap": "50ms", "join": join}, names)
    writes = []
    for i in range(3):
        for nm, m in (("flow_input1", m1), ("flow_input2", m2)):
            mb = bin_batch(m)
            mb.input_name = nm
            writes.append((nm, mb.record_batch))
            win.write(mb)
    done, r, _ = read_with_timeout(win, 2.0)
    assert done and r is not None
    got = r[0].record_batch
    want = window_drain(writes, join, names)
    assert got.schema.names == want.schema.names == ["id", "value", "sensor", "id", "value", "sensor"]
    assert got.num_rows == want.num_rows == 9
    assert sorted(map(repr, zip(*[c.to_pylist() for c in got.columns]))) == sorted(map(repr, zip(*[c.to_pylist() for c in want.columns])))


def test_join_window_missing_input_gives_empty_schema_batch(gpu):
    # join.rs:102-109
    join = {"query": "SELECT * FROM flow_input1 join flow_input2 on (flow_input1.id = flow_input2.id)", "codec": {"type": "json"}}
    win = TumblingWindow({"interval": "30ms", "join": join}, ["flow_input1", "flow_input2"])
    mb = bin_batch(b'{"id": 1}')
    mb.input_name = "flow_input1"
    win.write(mb)
    done, r, _ = read_with_timeout(win, 1.0)
    assert done and r is not None and r[0].num_rows == 0 and r[0].record_batch.num_columns == 0


def test_buffer_config_errors(gpu):
    for cls, msg in ((MemoryBuffer, "Memory buffer configuration is missing"), (SessionWindow, "Session window configuration is missing"),
                     (TumblingWindow, "Tumbling window configuration is missing")):
        with pytest.raises(ArkError) as e:
            cls(None)
        assert e.value.kind == "Config" and e.value.message == msg
    with pytest.raises(ArkError) as e:
        TumblingWindow({"size": "1s"})  # docs/…/join.md:62-64 writes `size`; the serde struct wants `interval`
    assert e.value.kind == "Serialization"
    with pytest.raises(ArkError):
        MemoryBuffer({"capacity": 10, "timeout": "soon"})
    for good in ("1s", "500ms", "100us", "1ns", "2m", "1h 30m"):
        MemoryBuffer({"capacity": 1, "timeout": good}).close()
