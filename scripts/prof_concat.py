import cProfile, pstats, sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import torch
from arkflow_b200 import _lib as L, arrow_ffi as F
from arkflow_b200.buffer import concat_batches_device
from arkflow_b200.processor import _check
lib = L.lib(); _check(lib.ark_b200_init(0))
n = 1 << 24
def synth(n, row0=0):
    dev, sch = L.ArrowDeviceArray(), L.ArrowSchema()
    _check(lib.ark_synth_batch_device(n, row0, 42, 0, 1000000, C.byref(dev), C.byref(sch)))
    return F.DeviceBatch.adopt(dev, sch)
parts = [synth(n // 16, i * (n // 16)) for i in range(16)]
for _ in range(3): concat_batches_device(parts).close()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): concat_batches_device(parts).close()
torch.cuda.synchronize()
print("per call ms", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): concat_batches_device(parts).close()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
lib.ark_kernel_timing_reset(); lib.ark_kernel_timing_enable(1)
t0 = time.perf_counter()
for _ in range(5): concat_batches_device(parts).close()
torch.cuda.synchronize()
print("per call ms (timing on)", (time.perf_counter() - t0) / 5 * 1e3)
ms, nn = C.c_double(), C.c_int64()
lib.ark_kernel_timing_get(b"concat_copy_kernel", C.byref(ms), C.byref(nn)); print("copy kernel ms", ms.value / max(nn.value, 1))
lib.ark_kernel_timing_enable(0)
# the same after a large unrelated allocation pattern (what bench_configs.py does before its concat leg)
from arkflow_b200.processor import SqlProcessor
big = [synth(n, i * n) for i in range(3)]
proc = SqlProcessor({"query": "SELECT sensor, SUM(value), COUNT(*) FROM flow GROUP BY sensor"})
for b in big: proc.process_device(b).close()
del big
t0 = time.perf_counter()
for _ in range(5): concat_batches_device(parts).close()
torch.cuda.synchronize()
print("per call ms (after other work)", (time.perf_counter() - t0) / 5 * 1e3)
parts2 = [synth(n // 16, i * (n // 16)) for i in range(16)]
for _ in range(2): concat_batches_device(parts2).close()
t0 = time.perf_counter()
for _ in range(5): concat_batches_device(parts2).close()
torch.cuda.synchronize()
print("per call ms (fresh parts)", (time.perf_counter() - t0) / 5 * 1e3)
