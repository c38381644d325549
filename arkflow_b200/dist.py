"""Multi-GPU execution of the sql processor: one process per GPU, torch.distributed for the plumbing.

SURVEY.md §8(e):
  * filter / project / json decode shard by rows — no collective (each rank processes its batches);
  * GROUP BY: each rank partial-aggregates its rows (hash_agg kernel), the partial states are
    hash-partitioned by key owner and exchanged — over peer memory on GPUs (exchange_partitions_p2p: every
    rank pulls its slices of its peers' batches with ONE segmented-copy kernel over NVLink), or with one
    all-to-all(v) per buffer (NCCL; gloo on CPU for the host-logic tests) — and the owner merges (same
    kernel in merge mode).  This mirrors DataFusion's
    AggregateExec(Partial) → RepartitionExec(Hash) → AggregateExec(FinalPartitioned), which the
    reference reaches in-process (crates/arkflow-plugin/src/processor/sql.rs:126-129);
  * JOIN: both sides are hash-partitioned on the join key and exchanged the same way, then joined
    locally (buffer/join.rs:111-118 runs the join on one node).

The exchange code only touches torch tensors, so it is device-agnostic; the compute steps go through
an `engine` object — `NativeEngine` (C ABI → CUDA kernels) in production.  Tests substitute a CPU
engine to exercise this host logic under gloo with world_size 2.
"""
from __future__ import annotations

import os
import sys
import time
import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib as L
from .arrow_ffi import DeviceBatch, DeviceColumn


class NativeEngine:
    """Compute steps on the local GPU through the C ABI."""

    def __init__(self, query: str, table_name: str = "flow"):
        from .processor import SqlProcessor

        self.proc = SqlProcessor({"query": query, "table_name": table_name})

    def partial_aggregate(self, batch: DeviceBatch, n_parts: int):
        from .processor import _check

        lib = L.lib()
        dev, sch = batch.export()
        out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
        rows = (C.c_int64 * n_parts)()
        try:
            status = lib.ark_sql_partial_aggregate_device(self.proc._h, C.byref(dev), C.byref(sch), n_parts, C.byref(out_dev), C.byref(out_sch), rows)
        finally:
            from .arrow_ffi import release_array, release_schema

            release_schema(sch)
            release_array(dev.array)
        _check(status)
        return DeviceBatch.adopt(out_dev, out_sch), list(rows)

    def final_aggregate(self, partial: DeviceBatch) -> DeviceBatch:
        from .processor import _check

        lib = L.lib()
        dev, sch = partial.export()
        out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
        try:
            status = lib.ark_sql_final_aggregate_device(self.proc._h, C.byref(dev), C.byref(sch), C.byref(out_dev), C.byref(out_sch))
        finally:
            from .arrow_ffi import release_array, release_schema

            release_schema(sch)
            release_array(dev.array)
        _check(status)
        return DeviceBatch.adopt(out_dev, out_sch)

    def process(self, batch: DeviceBatch) -> Optional[DeviceBatch]:
        return self.proc.process_device(batch)

    # ---- device-side exchange (csrc/group_exchange.cu) ----
    def group_by_push(self, batch: DeviceBatch, ctx: "ExchangeContext") -> None:
        from .arrow_ffi import release_array, release_schema
        from .processor import _check

        dev, sch = batch.export()
        try:
            status = L.lib().ark_sql_group_by_push_device(self.proc._h, ctx._h, C.byref(dev), C.byref(sch))
        finally:
            release_schema(sch)
            release_array(dev.array)
        _check(status)

    def group_by_merge(self, ctx: "ExchangeContext") -> Optional[DeviceBatch]:
        """This rank's share of the groups, or None when some rank held a key that cannot travel inline (every rank
        gets None for that step and falls back to the descriptor exchange)."""
        from .processor import _check

        out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
        status = L.lib().ark_sql_group_by_merge_device(self.proc._h, ctx._h, C.byref(out_dev), C.byref(out_sch))
        if status == L.ARK_ERR_UNSUPPORTED:
            return None
        _check(status)
        return DeviceBatch.adopt(out_dev, out_sch)

    def group_by_exchange(self, batch: DeviceBatch, ctx: "ExchangeContext") -> Optional[DeviceBatch]:
        self.group_by_push(batch, ctx)
        return self.group_by_merge(ctx)

    def hash_partition(self, batch: DeviceBatch, key_column: str, n_parts: int):
        from .arrow_ffi import release_array, release_schema
        from .processor import _check

        lib = L.lib()
        dev, sch = batch.export()
        out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
        rows = (C.c_int64 * n_parts)()
        try:
            status = lib.ark_hash_partition_device(C.byref(dev), C.byref(sch), key_column.encode(), n_parts, C.byref(out_dev), C.byref(out_sch), rows)
        finally:
            release_schema(sch)
            release_array(dev.array)
        _check(status)
        return DeviceBatch.adopt(out_dev, out_sch), list(rows)

    def join(self, tables: dict) -> DeviceBatch:
        return self.proc.process_tables_device(tables)


class ExchangeContext:
    """One rank's end of the device-side GROUP BY exchange (include/arkflow_b200.h: ark_dist_*).  `region_bytes` must
    hold the partial states one source sends this rank in one step (32 bytes per group for up to two accumulators)."""

    def __init__(self, rank: int, world: int, region_bytes: int):
        from .processor import _check

        self.rank, self.world, self.region_bytes = rank, world, region_bytes
        self._h = C.c_void_p()
        _check(L.lib().ark_dist_create(rank, world, region_bytes, C.byref(self._h)))

    def handle(self) -> bytes:
        from .processor import _check

        n = int(L.lib().ark_dist_handle_bytes())
        blob = (C.c_uint8 * n)()
        size = C.c_int64(0)
        _check(L.lib().ark_dist_export(self._h, blob, n, C.byref(size)))
        return bytes(blob[: size.value])

    def connect(self, handles: list) -> None:
        from .processor import _check

        assert len(handles) == self.world
        stride = len(handles[0])
        buf = (C.c_uint8 * (stride * self.world)).from_buffer_copy(b"".join(handles))
        _check(L.lib().ark_dist_connect(self._h, buf, stride))

    @classmethod
    def from_process_group(cls, region_bytes: int, group=None) -> "ExchangeContext":
        """One context per rank of a torch.distributed group; the handles travel in one all_gather."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ctx = cls(rank, world, region_bytes)
        mine = torch.frombuffer(bytearray(ctx.handle()), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            mine = mine.cuda()
        everyone = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine, group=group)
        ctx.connect([bytes(t.cpu().numpy()) for t in everyone])
        dist.barrier(group)
        return ctx

    def stats(self) -> dict:
        out = (C.c_int64 * 4)()
        L.lib().ark_dist_stats(self._h, out)
        return {"steps": out[0], "records_received": out[1], "groups": out[2], "region_bytes": out[3]}

    def close(self) -> None:
        if self._h:
            L.lib().ark_dist_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# all-to-all(v) of a partition-ordered batch
# ---------------------------------------------------------------------------------------------
def _all_to_all_rows(t: torch.Tensor, send_counts: list[int], recv_counts: list[int], group) -> torch.Tensor:
    out = torch.empty(sum(recv_counts), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out, t.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    return out


def _unpack_bits(bits: torch.Tensor, n: int) -> torch.Tensor:
    idx = torch.arange(n, device=bits.device)
    return ((bits[idx >> 3].to(torch.int32) >> (idx & 7).to(torch.int32)) & 1).to(torch.uint8)


def _pack_bits(bytes_: torch.Tensor) -> torch.Tensor:
    n = bytes_.numel()
    pad = (-n) % 8
    b = torch.cat([bytes_, torch.zeros(pad, dtype=torch.uint8, device=bytes_.device)]).view(-1, 8).to(torch.int32)
    w = (b << torch.arange(8, device=bytes_.device, dtype=torch.int32)).sum(dim=1)
    return w.to(torch.uint8)


def exchange_partitions(batch: DeviceBatch, part_rows: list[int], group=None) -> DeviceBatch:
    """Rows [sum(part_rows[:p]), sum(part_rows[:p+1])) of `batch` go to rank p; returns the rows this
    rank received (in source-rank order).  One all-to-all(v) per buffer + one for the counts."""
    world = dist.get_world_size(group)
    assert len(part_rows) == world
    device = batch.columns[0].data.device if batch.columns else torch.device("cpu")
    send = torch.tensor(part_rows, dtype=torch.int64, device=device)
    recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, group=group)
    send_counts, recv_counts = [int(x) for x in part_rows], [int(x) for x in recv.tolist()]
    n_out = sum(recv_counts)
    row_starts = [0]
    for c in send_counts:
        row_starts.append(row_starts[-1] + c)
    cols = []
    for c in batch.columns:
        validity = None
        if c.validity is not None and c.validity.numel():
            vb = _all_to_all_rows(_unpack_bits(c.validity, c.length), send_counts, recv_counts, group)
            validity = _pack_bits(vb)
        if c.dtype in ("int64", "float64"):
            data = _all_to_all_rows(c.data[: c.length], send_counts, recv_counts, group)
            cols.append(DeviceColumn(c.name, c.dtype, n_out, data, None, validity, -1 if validity is not None else 0, c.nullable))
        elif c.dtype == "bool":
            data = _pack_bits(_all_to_all_rows(_unpack_bits(c.data, c.length), send_counts, recv_counts, group))
            cols.append(DeviceColumn(c.name, c.dtype, n_out, data, None, validity, -1 if validity is not None else 0, c.nullable))
        else:
            offs = c.offsets[: c.length + 1].to(torch.int64)
            lens = (offs[1:] - offs[:-1]).to(torch.int32)
            rlens = _all_to_all_rows(lens, send_counts, recv_counts, group)
            bounds = offs[torch.tensor(row_starts, device=device)]
            bsend = (bounds[1:] - bounds[:-1]).to(torch.int64)
            brecv = torch.empty(world, dtype=torch.int64, device=device)
            dist.all_to_all_single(brecv, bsend, group=group)
            first = int(offs[0].item()) if c.length else 0
            payload = c.data[first: first + int(bsend.sum().item())]
            rbytes = _all_to_all_rows(payload, [int(x) for x in bsend.tolist()], [int(x) for x in brecv.tolist()], group)
            roffs = torch.zeros(n_out + 1, dtype=torch.int32, device=device)
            if n_out:
                roffs[1:] = torch.cumsum(rlens.to(torch.int64), 0).to(torch.int32)
            cols.append(DeviceColumn(c.name, c.dtype, n_out, rbytes, roffs, validity, -1 if validity is not None else 0, c.nullable))
    return DeviceBatch(cols, n_out)


_IPC_SLOT_BYTES = 8192  # descriptor bytes per rank in the exchange's all-gather (24 + 400 per column)


def _ipc_export(batch: DeviceBatch) -> Optional[bytes]:
    """CUDA-IPC descriptor of a device batch, or None when its memory cannot be exported."""
    from .arrow_ffi import release_array, release_schema

    lib = L.lib()
    cap = 96 + 400 * max(len(batch.columns), 1)
    blob = (C.c_uint8 * cap)()
    size = C.c_int64(0)
    dev, sch = batch.export()
    try:
        status = lib.ark_ipc_export_device(C.byref(dev), C.byref(sch), blob, cap, C.byref(size))
    finally:
        release_schema(sch)
        release_array(dev.array)
    return bytes(blob[: size.value]) if status == 0 else None


def exchange_partitions_p2p(batch: DeviceBatch, part_rows: list[int], group=None) -> Optional[DeviceBatch]:
    """The same exchange as exchange_partitions, over peer memory (csrc/ipc_exchange.cu): every rank publishes
    its partition-ordered batch as CUDA IPC handles and PULLS its slice of every peer's batch with one
    segmented-copy launch — NVLink transfer, concatenation and offset rebasing in the same kernel.
    Returns None (on every rank alike) when some rank's memory cannot be exported; callers then use NCCL."""
    from .processor import _check

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    trace = os.environ.get("ARK_DIST_TRACE") and rank == 0
    t_ = [time.perf_counter()]

    def lap():
        if trace:
            torch.cuda.synchronize()
            t_.append(time.perf_counter())

    blob = _ipc_export(batch)
    lap()
    # ONE fixed-size all-gather carries every rank's descriptor and partition row counts:
    # [ok:int64 | blob_len:int64 | part_rows: world × int64 | blob bytes, zero-padded]
    slot = 16 + 8 * world + _IPC_SLOT_BYTES
    if blob is not None and len(blob) > _IPC_SLOT_BYTES:
        blob = None
    head = [0 if blob is None else 1, 0 if blob is None else len(blob)] + [int(x) for x in part_rows]
    mine = torch.zeros(slot, dtype=torch.uint8)
    mine[: 16 + 8 * world] = torch.tensor(head, dtype=torch.int64).view(torch.uint8)
    if blob is not None:
        mine[16 + 8 * world: 16 + 8 * world + len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    device = torch.device("cuda", torch.cuda.current_device())
    everyone = torch.empty(slot * world, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(everyone, mine.to(device), group=group)
    everyone = everyone.cpu()
    lap()
    gathered = []
    for s_ in range(world):
        rec = everyone[s_ * slot: (s_ + 1) * slot]
        h = rec[: 16 + 8 * world].view(torch.int64).tolist()
        gathered.append((bytes(rec[16 + 8 * world: 16 + 8 * world + h[1]].numpy()) if h[0] else None, h[2:]))
    if any(g[0] is None for g in gathered):
        return None
    blobs = [g[0] for g in gathered]
    row0 = (C.c_int64 * world)(*[sum(g[1][:rank]) for g in gathered])
    nrows = (C.c_int64 * world)(*[g[1][rank] for g in gathered])
    keep = [(C.c_uint8 * len(b)).from_buffer_copy(b) for b in blobs]
    ptrs = (C.POINTER(C.c_uint8) * world)(*[C.cast(k, C.POINTER(C.c_uint8)) for k in keep])
    sizes = (C.c_int64 * world)(*[len(b) for b in blobs])
    out_dev, out_sch = L.ArrowDeviceArray(), L.ArrowSchema()
    lap()
    status = L.lib().ark_ipc_concat_slices_device(world, ptrs, sizes, row0, nrows, C.byref(out_dev), C.byref(out_sch))
    lap()
    dist.barrier(group)  # every reader is done with this rank's buffers before they are released
    lap()
    if trace:
        d = [(b - a) * 1e3 for a, b in zip(t_, t_[1:])]
        print("[exchange p2p] rows %d: export %.2f ms, all-gather %.2f, decode %.2f, pull+concat %.2f, barrier %.2f" % (batch.num_rows, *d), file=sys.stderr)
    _check(status)
    return DeviceBatch.adopt(out_dev, out_sch)


def _exchange(batch: DeviceBatch, part_rows: list[int], group, p2p: Optional[bool]) -> DeviceBatch:
    """p2p: True / False force the path; None = peer memory when the backend is NCCL (GPU ranks), else the
    torch all-to-all (gloo on CPU)."""
    import os

    if p2p is None:
        p2p = dist.get_backend(group) == "nccl" and os.environ.get("ARK_DIST_EXCHANGE", "p2p") != "nccl"
    if p2p:
        out = exchange_partitions_p2p(batch, part_rows, group)
        if out is not None:
            return out
    return exchange_partitions(batch, part_rows, group)


def distributed_group_by(engine, local_batch: DeviceBatch, group=None, p2p: Optional[bool] = None,
                         ctx: Optional[ExchangeContext] = None) -> DeviceBatch:
    """GROUP BY over the union of every rank's `local_batch`; returns this rank's share of the groups
    (group owners are disjoint, so the concatenation over ranks is the full result).  With an ExchangeContext the
    partial states are pushed over NVLink by the kernel that scans the partial table and merged by the owner without
    a host round trip (csrc/group_exchange.cu); keys that cannot travel inline make every rank fall back, for that
    batch, to the descriptor exchange below."""
    world = dist.get_world_size(group)
    if ctx is not None:
        out = engine.group_by_exchange(local_batch, ctx)
        if out is not None:
            return out
    partial, part_rows = engine.partial_aggregate(local_batch, world)
    keyless = not partial.columns or partial.columns[0].name.startswith("__acc")
    received = _exchange(partial, part_rows, group, p2p)
    partial.close()
    result = engine.final_aggregate(received)
    received.close()
    if keyless and dist.get_rank(group) != 0:
        # a global aggregate has one group, owned by rank 0; other ranks merged nothing
        result = DeviceBatch([DeviceColumn(c.name, c.dtype, 0, c.data[:0], None if c.offsets is None else c.offsets[:1], None, 0, c.nullable)
                              for c in result.columns], 0)
    return result


def distributed_join(engine, tables: dict, keys: dict, group=None, p2p: Optional[bool] = None) -> DeviceBatch:
    """Inner equi-join over the union of every rank's tables.  `keys[name]` is the join column of table
    `name`.  Both sides are hash-partitioned on the key and exchanged with one all-to-all(v) each
    (SURVEY.md §8(e)); the local join then sees every row of its key range.  Output stays sharded."""
    world = dist.get_world_size(group)
    local = {}
    for name, batch in tables.items():
        parted, rows = engine.hash_partition(batch, keys[name], world)
        local[name] = _exchange(parted, rows, group, p2p)
        # released here, not whenever the garbage collector gets to it: the arena the partition was published from is
        # reusable as soon as its buffers are gone, and a rank that publishes from a NEW arena costs every peer a
        # cudaIpcOpenMemHandle (the exchange has ended with a barrier: no peer reads it any more)
        parted.close()
    out = engine.join(local)
    for b in local.values():
        b.close()
    return out
