"""torchrun entry: phase times of the distributed join / GROUP BY of bench.py's sharded workloads (host wall clock with a
device synchronize after every phase, so the phases add up to more than the pipelined step)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from arkflow_b200 import _lib as L
    import arkflow_b200.dist as D
    from arkflow_b200.arrow_ffi import DeviceBatch
    from arkflow_b200.processor import _check
    lib = L.lib()
    _check(lib.ark_b200_init(local))

    def synth(n, row0, K):
        a, s = L.ArrowDeviceArray(), L.ArrowSchema()
        _check(lib.ark_synth_batch_device(n, row0, 3, 0, K, C.byref(a), C.byref(s)))
        return DeviceBatch.adopt(a, s)

    n_probe, n_build = 1 << 24, 1 << 20
    K = n_build * world
    eng = D.NativeEngine("SELECT * FROM flow_input1 JOIN flow_input2 ON flow_input1.sensor = flow_input2.sensor")
    probe, build = synth(n_probe, (1 << 42) + rank * (1 << 34), K), synth(n_build, (1 << 43) + rank * n_build, K)
    T = {}

    def tick(name, t0):
        torch.cuda.synchronize(); T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    real_export, real_concat = D._ipc_export, None
    for it in range(8):
        dist.barrier(); torch.cuda.synchronize()
        t_all = time.perf_counter()
        t0 = time.perf_counter(); pp, prow = eng.hash_partition(probe, "sensor", world); tick("partition probe", t0)
        t0 = time.perf_counter(); pb, brow = eng.hash_partition(build, "sensor", world); tick("partition build", t0)
        t0 = time.perf_counter(); xp = D._exchange(pp, prow, None, True); tick("exchange probe", t0)
        t0 = time.perf_counter(); xb = D._exchange(pb, brow, None, True); tick("exchange build", t0)
        t0 = time.perf_counter(); out = eng.join({"flow_input1": xp, "flow_input2": xb}); tick("local join", t0)
        tick("step", t_all)
        rows = out.num_rows
        del pp, pb, xp, xb; out.close()
    if rank == 0:
        print(f"world {world}, output rows/rank {rows}")
        for k, v in T.items():
            print(f"{k:18s} first {v[0]:8.2f} ms   steady (median of last 5) {sorted(v[-5:])[2]:8.2f} ms")
    dist.destroy_process_group()


main()
