"""The CUDA path against the reference's own pins (tests/golden/reference_pins.json)."""
import pytest

from arkflow_b200.processor import MessageBatch, SqlProcessor
from golden_util import check_expect, load_pins, pin_batch

pytestmark = pytest.mark.gpu
PINS = [p for p in load_pins() if p["expect"]["kind"] != "ConstructError"]


@pytest.mark.parametrize("pin", PINS, ids=[p["id"] for p in PINS])
def test_library_matches_reference_pin(gpu, pin):
    cfg = {"query": pin["query"]}
    if "table_name" in pin:
        cfg["table_name"] = pin["table_name"]
    proc = SqlProcessor(cfg)
    rb = pin_batch(pin)
    for _ in range(pin.get("repeat", 1)):
        r = proc.process(MessageBatch.new_arrow(rb))
        check_expect(pin, None if r.is_none() else r.batches[0].record_batch)
