#!/usr/bin/env python
"""bench.py — rows/s through the `sql` processor on synthetic Arrow batches (BASELINE.json metric).

Workload (BASELINE.json configs[1]): `SELECT sensor, value FROM flow WHERE value >= 10` over a
2^30-row table of schema S (timestamp Int64, value Int64, sensor Utf8 "temp_%07d"), fed as 64
RecordBatches of 2^24 rows (one Utf8 array cannot exceed 2 GiB).  A *step* is one process() call on
one 2^24-row batch; the default 64 steps are the whole 2^30-row job.  N>1: every rank owns its own
2^30-row shard (weak scaling, no collective on this path — SURVEY.md §8(e)).

  value   device-resident: ark_sql_process_device on batches already in HBM
  e2e     reference-facing call with HOST buffers: ark_sql_process on pinned host Arrow buffers,
          H2D + kernel + D2H inside the timed region
  roofline  filter_project_kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port (numpy/pyarrow restatement; the Rust reference cannot be built here)

`--impl reference` times that oracle port on the host cores with the same JSON contract.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

QUERY = "SELECT sensor, value FROM flow WHERE value >= 10"
ROWS_PER_BATCH = 1 << 24
N_BATCHES = 64  # × 2^24 = 2^30 rows per GPU
SEED = 42
KEY_SPACE = 1_000_000
METRIC = "rows/sec through sql processor on synthetic Arrow batches"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
# CPU side: oracle port, all host cores (partition-parallel like DataFusion's target_partitions)
# ------------------------------------------------------------------------------------------------
def cpu_process_batch(rb, pool, n_threads, chunk_rows=1 << 20):
    from oracle.sql_oracle import sql_process

    chunks = [rb.slice(o, min(chunk_rows, rb.num_rows - o)) for o in range(0, rb.num_rows, chunk_rows)]
    outs = list(pool.map(lambda c: sql_process(c, QUERY), chunks))
    return sum(o.num_rows for o in outs if o is not None)


def cpu_baseline_measure(sample_rows, budget_s=12.0):
    from oracle.synth import synth_batch

    n_threads = os.cpu_count() or 1
    rb = synth_batch(sample_rows, seed=SEED, key_space=KEY_SPACE)
    with ThreadPoolExecutor(max_workers=n_threads) as pool:
        cpu_process_batch(rb, pool, n_threads)  # warm-up
        t0 = time.perf_counter()
        passes = 0
        while True:
            cpu_process_batch(rb, pool, n_threads)
            passes += 1
            if time.perf_counter() - t0 >= budget_s or passes >= 50:
                break
        dt = time.perf_counter() - t0
    return {"value": passes * sample_rows / dt, "unit": "rows/s", "cores": n_threads, "kind": "port",
            "sample": f"{passes} passes over a {sample_rows}-row batch of the same workload (oracle port: numpy/pyarrow, "
                      f"{n_threads} threads over 2^20-row partitions); CPU proxy, not DataFusion"}


def run_reference(args):
    """The reference's CPU implementation of the path = the oracle port (the Rust reference cannot be
    compiled in this image).  Each step = one bounded 2^22-row batch of the workload."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle.synth import synth_batch

    sample_rows = 1 << 22
    n_threads = os.cpu_count() or 1
    batches = [synth_batch(sample_rows, row0=i * sample_rows, seed=SEED, key_space=KEY_SPACE) for i in range(4)]
    with ThreadPoolExecutor(max_workers=n_threads) as pool:
        for i in range(args.warmup):
            cpu_process_batch(batches[i % 4], pool, n_threads)
        t0 = time.perf_counter()
        for i in range(args.steps):
            cpu_process_batch(batches[i % 4], pool, n_threads)
        dt = time.perf_counter() - t0
    value = args.steps * sample_rows / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "filter+project SELECT sensor,value WHERE value>=10 (BASELINE configs[1])",
                   "rows_per_step": sample_rows, "query": QUERY, "schema": "timestamp:Int64,value:Int64,sensor:Utf8(12B)"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": n_threads, "kind": "port",
                         "sample": f"{args.steps} steps of a {sample_rows}-row batch; oracle port (numpy/pyarrow), {n_threads} threads"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        if os.environ.get("ARK_BENCH_NO_SAMPLER"):  # experiment knob: how much does polling nvidia-smi cost the timed region?
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from arkflow_b200 import _lib as L
    from arkflow_b200 import arrow_ffi as F
    from arkflow_b200.processor import SqlProcessor, _check
    import pyarrow as pa

    lib = L.lib()
    _check(lib.ark_b200_init(local_rank))
    proc = SqlProcessor({"query": QUERY})
    shard_row0 = rank * N_BATCHES * ROWS_PER_BATCH

    # ---- resident inputs: 64 batches of 2^24 rows generated in HBM ----
    n_resident = min(N_BATCHES, max(args.steps, 8))
    resident = []
    for b in range(n_resident):
        dev, sch = L.ArrowDeviceArray(), L.ArrowSchema()
        _check(lib.ark_synth_batch_device(ROWS_PER_BATCH, shard_row0 + b * ROWS_PER_BATCH, SEED, 0, KEY_SPACE, C.byref(dev), C.byref(sch)))
        resident.append(F.DeviceBatch.adopt(dev, sch))
    for b in resident:  # build each resident batch's Arrow C struct tree now: describing the inputs is set-up, not a step
        d, s_ = b.export()
        F.release_schema(s_)
        F.release_array(d.array)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timed region (single caller: one process() after another) ----
    def device_step(i):
        out = proc.process_device(resident[i % n_resident])
        rows = out.num_rows
        out.close()
        return rows

    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(local_rank)
    lib.ark_kernel_timing_reset()
    lib.ark_kernel_timing_enable(1)
    barrier()
    sampler.start()
    launches0 = lib.ark_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    out_rows = 0
    for i in range(args.steps):
        out_rows += device_step(i)
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    dev_ms_local = ev0.elapsed_time(ev1)
    launches = lib.ark_kernel_launch_count() - launches0
    barrier()
    lib.ark_kernel_timing_enable(0)
    kms, kn = C.c_double(), C.c_int64()
    lib.ark_kernel_timing_get(b"filter_project_tma_kernel", C.byref(kms), C.byref(kn))
    dev_ms = max_over_ranks(dev_ms_local)
    # extra: the same K steps driven by several host threads, as the reference's `thread_num` workers do
    # (crates/arkflow-core/src/stream/mod.rs:117-126; process() is re-entrant).  Not the contract `value`.
    dthreads = max(1, args.device_threads)
    conc = None
    if dthreads > 1:
        with ThreadPoolExecutor(max_workers=dthreads) as dpool:
            work = lambda t, hi: sum(device_step(i) for i in range(t, hi, dthreads))
            list(dpool.map(lambda t: work(t, max(args.warmup, dthreads)), range(dthreads)))
            barrier()
            ev0.record()
            list(dpool.map(lambda t: work(t, args.steps), range(dthreads)))
            torch.cuda.synchronize()
            ev1.record()
            ev1.synchronize()
        conc_ms = max_over_ranks(ev0.elapsed_time(ev1))
        conc = {"value": args.steps * ROWS_PER_BATCH * world / (conc_ms / 1e3), "unit": "rows/s", "host_threads": dthreads,
                "ms_per_step": conc_ms / args.steps}
    rows_total = args.steps * ROWS_PER_BATCH * world
    value = rows_total / (dev_ms / 1e3)

    # roofline of the dominant kernel: algorithmic bytes of one launch = n·(8 value + 4 offsets + 12 bytes)
    # read + kept·(8 + 4 + 12) written (36 B/row at σ = 0.5; exact σ from the run)
    kept = out_rows / max(args.steps, 1)
    alg_bytes = ROWS_PER_BATCH * 24 + kept * 24 + 4
    peak, peak_src = load_peaks()
    k_avg_ms = kms.value / max(kn.value, 1)
    achieved = alg_bytes / (k_avg_ms / 1e3) / 1e9 if k_avg_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_filter_project_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- e2e: host (pinned) Arrow buffers in, host Arrow buffers out, copies inside the timed region ----
    n_host = 4
    host_batches, keep = [], []
    for b in range(n_host):
        db = resident[b]
        arrays = []
        for c in db.columns:
            if c.dtype == "int64":
                t = torch.empty(c.length, dtype=torch.int64, pin_memory=True)
                t.copy_(c.data)
                keep.append(t)
                arrays.append(pa.Array.from_buffers(pa.int64(), c.length, [None, pa.foreign_buffer(t.data_ptr(), c.length * 8, base=t)]))
            else:
                to = torch.empty(c.length + 1, dtype=torch.int32, pin_memory=True)
                to.copy_(c.offsets)
                dsrc = c.data.tensor() if hasattr(c.data, "tensor") else c.data
                td = torch.empty(dsrc.numel(), dtype=torch.uint8, pin_memory=True)
                td.copy_(dsrc)
                keep += [to, td]
                arrays.append(pa.Array.from_buffers(pa.utf8(), c.length, [None, pa.foreign_buffer(to.data_ptr(), (c.length + 1) * 4, base=to),
                                                                          pa.foreign_buffer(td.data_ptr(), td.numel(), base=td)]))
        torch.cuda.synchronize()
        host_batches.append(pa.RecordBatch.from_arrays(arrays, names=[c.name for c in db.columns]))
    h2d_per_step = ROWS_PER_BATCH * 8 + (ROWS_PER_BATCH + 1) * 4 + ROWS_PER_BATCH * 12
    e2e_threads = args.e2e_threads
    d2h_acc = [0] * e2e_threads

    def e2e_worker(tid, steps):
        for i in steps:
            r = proc.process(host_batches[i % n_host])
            rb = r.batches[0].record_batch
            d2h_acc[tid] += rb.nbytes
            del r, rb

    e2e_steps = max(args.e2e_steps, e2e_threads)
    with ThreadPoolExecutor(max_workers=e2e_threads) as pool:
        list(pool.map(lambda t: e2e_worker(t, range(t, max(args.warmup, e2e_threads), e2e_threads)), range(e2e_threads)))
        d2h_acc[:] = [0] * e2e_threads
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        list(pool.map(lambda t: e2e_worker(t, range(t, e2e_steps, e2e_threads)), range(e2e_threads)))
        torch.cuda.synchronize()
        ev1.record()
        ev1.synchronize()
        e2e_ms_local = ev0.elapsed_time(ev1)
        e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()  # sampled across both timed regions (device-resident and end-to-end)
    e2e_ms = max_over_ranks(max(e2e_ms_local, e2e_wall_ms))
    e2e_value = e2e_steps * ROWS_PER_BATCH * world / (e2e_ms / 1e3)
    d2h_per_step = sum(d2h_acc) / e2e_steps

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_measure(1 << 22)
    barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "filter+project SELECT sensor,value WHERE value>=10 on 2^30-row int64/Utf8 table, 64 batches of 2^24 rows per GPU (BASELINE configs[1])",
                       "query": QUERY, "rows_per_step": ROWS_PER_BATCH, "resident_batches": n_resident,
                       "schema": "timestamp:Int64,value:Int64,sensor:Utf8(12B)", "selectivity": kept / ROWS_PER_BATCH,
                       "l2": "inputs larger than L2 (537 MB per batch, distinct batch each step)", "parallelism": f"{world} rank(s), row shards, no collective"},
            "roofline": {"bound": "hbm", "kernel": "filter_project_tma_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_avg_ms, "launches_timed": kn.value},
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h_per_step,
                    "steps": e2e_steps, "host_threads": e2e_threads, "ms_per_step": e2e_ms / e2e_steps,
                    "note": "pinned host Arrow buffers → ark_sql_process → host Arrow buffers; PCIe-bound"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if conc is not None:
            line["concurrent_callers"] = conc
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=16)
    ap.add_argument("--e2e-threads", type=int, default=3)
    ap.add_argument("--device-threads", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
