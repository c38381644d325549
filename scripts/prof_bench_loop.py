"""cProfile of bench.py's device-resident loop (same objects, same calls), to see where host time goes."""
import cProfile, pstats, sys, os, ctypes as C, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from arkflow_b200 import _lib as L, arrow_ffi as F
from arkflow_b200.processor import SqlProcessor, _check
lib = L.lib(); _check(lib.ark_b200_init(0))
proc = SqlProcessor({"query": bench.QUERY})
resident = []
for b in range(16):
    dev, sch = L.ArrowDeviceArray(), L.ArrowSchema()
    _check(lib.ark_synth_batch_device(bench.ROWS_PER_BATCH, b * bench.ROWS_PER_BATCH, bench.SEED, 0, bench.KEY_SPACE, C.byref(dev), C.byref(sch)))
    resident.append(F.DeviceBatch.adopt(dev, sch))
def device_step(i):
    out = proc.process_device(resident[i % 16])
    rows = out.num_rows
    out.close()
    return rows
for i in range(8): device_step(i)
for timing in (0, 1):
    lib.ark_kernel_timing_reset(); lib.ark_kernel_timing_enable(timing)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(64): device_step(i)
    torch.cuda.synchronize(); print(f"timing={timing}: per step {(time.perf_counter()-t0)/64*1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for i in range(64): device_step(i)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
