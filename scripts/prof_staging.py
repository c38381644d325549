"""Host memcpy bandwidth of this box: N threads copying 4 MB chunks from pageable numpy memory into (a) pageable, (b) pinned
memory — the staging step of ark_sql_process with pageable inputs (csrc/batch.cu: staged_h2d)."""
import sys, time, threading
import numpy as np
import torch

CH = 4 << 20
src = np.random.randint(0, 255, 512 << 20, dtype=np.uint8)
dst_page = np.empty(256 << 20, dtype=np.uint8); dst_page[:] = 0
pin = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True); pin.zero_(); dst_pin = pin.numpy()
for name, dst in (("pageable->pageable", dst_page), ("pageable->pinned", dst_pin)):
    for T in (1, 2, 4, 8, 16, 32):
        n_chunks = len(src) // CH
        def work(w):
            for c in range(w, n_chunks, T):
                o = (c * CH) % len(dst)
                np.copyto(dst[o:o + CH], src[c * CH:(c + 1) * CH])
        best = 0
        for rep in range(3):
            ths = [threading.Thread(target=work, args=(w,)) for w in range(T)]
            t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
            best = max(best, len(src) / dt / 1e9)
        print(f"{name} threads={T}: {best:.1f} GB/s", flush=True)
# plain H2D from pinned, one big copy vs 4 MB chunks
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for label, chunks in (("one 256 MB copy", 1), ("64 x 4 MB copies", 64)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(4):
        step = len(pin) // chunks
        for c in range(chunks):
            d[c * step:(c + 1) * step].copy_(pin[c * step:(c + 1) * step], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D pinned {label}: {4 * len(pin) / dt / 1e9:.1f} GB/s")
