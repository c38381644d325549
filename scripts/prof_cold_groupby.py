import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from arkflow_b200 import _lib as L, arrow_ffi as F
from arkflow_b200.processor import SqlProcessor, _check
lib = L.lib(); _check(lib.ark_b200_init(0))
dev, sch = L.ArrowDeviceArray(), L.ArrowSchema()
_check(lib.ark_synth_batch_device(1 << 24, 0, 42, 0, 1000000, C.byref(dev), C.byref(sch)))
b = F.DeviceBatch.adopt(dev, sch)
p = SqlProcessor({"query": "SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor"})
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = p.process_device(b); n = out.num_rows; out.close()
    print(f"call {i}: {1e3*(time.perf_counter()-t0):.2f} ms, groups {n}")
