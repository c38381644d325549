#!/usr/bin/env python
"""json_to_arrow on 2^22 device-resident 63-byte messages, repeatedly (for ncu / quick timing)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arkflow_b200 import _lib as L, arrow_ffi as F
from arkflow_b200.processor import JsonToArrowProcessor, _check
lib = L.lib(); _check(lib.ark_b200_init(0))
m = 1 << 22
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
msg = b'{ "timestamp": 1625000000000, "value": 10, "sensor": "temp_1" }'
data = torch.from_numpy(np.frombuffer(msg * m, dtype=np.uint8).copy()).cuda()
offs = torch.arange(0, (m + 1) * len(msg), len(msg), dtype=torch.int32, device="cuda")
payload = F.DeviceBatch([F.DeviceColumn("__value__", "binary", m, data, offs, None, 0, False)], m)
proc = JsonToArrowProcessor({})
for _ in range(3): proc.process_device(payload).close()
lib.ark_kernel_timing_reset(); lib.ark_kernel_timing_enable(1)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps): proc.process_device(payload).close()
torch.cuda.synchronize(); print(f"call wall avg {(time.perf_counter()-t0)/reps*1e3:.3f} ms")
for name in (b"json_parse_kernel", b"json_count_kernel", b"json_strings_kernel", b"pack_bits_kernel"):
    ms, n = C.c_double(), C.c_int64()
    lib.ark_kernel_timing_get(name, C.byref(ms), C.byref(n))
    if n.value: print(f"{name.decode():24s} avg {ms.value/n.value:.4f} ms over {n.value} launches")
