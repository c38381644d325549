import cProfile, pstats, sys, os, ctypes as C, time
sys.path.insert(0, os.getcwd())
import numpy as np, pyarrow as pa, torch
from arkflow_b200 import _lib as L, arrow_ffi as F
from arkflow_b200.buffer import concat_batches_device
from arkflow_b200.processor import _check, SqlProcessor
lib = L.lib(); _check(lib.ark_b200_init(0))
n = 1 << 24
def synth(n, row0=0, keys=1000000):
    dev, sch = L.ArrowDeviceArray(), L.ArrowSchema()
    _check(lib.ark_synth_batch_device(n, row0, 42, 0, keys, C.byref(dev), C.byref(sch)))
    return F.DeviceBatch.adopt(dev, sch)
K = 1000000
probe = synth(n)
bkeys = pa.array(["temp_%07d" % i for i in np.random.default_rng(0).permutation(K)])
build = F.DeviceBatch.from_arrow(pa.record_batch({"sensor": bkeys, "w": pa.array(np.arange(K), pa.int64())}))
jp = SqlProcessor({"query": "SELECT * FROM p JOIN b ON p.sensor = b.sensor"})
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    jp.process_tables_device({"p": probe, "b": build}).close()
del probe, build
parts = [synth(n // 16, i * (n // 16)) for i in range(16)]
torch.cuda.synchronize()
for i in range(14):
    t0 = time.perf_counter()
    o = concat_batches_device(parts)
    t1 = time.perf_counter()
    o.close()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"call {i}: concat {1e3*(t1-t0):.2f} ms, close {1e3*(t2-t1):.2f} ms, sync {1e3*(t3-t2):.2f} ms")
