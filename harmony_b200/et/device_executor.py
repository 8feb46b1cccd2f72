"""GPU executor: dense tables live in HBM slabs, everything else behaves like the host executor.

A `DeviceExecutor` owns ONE GPU (its CUDA context, a PeerArena, a side stream for migration).
Dense tables with `backend == "device"` get `DeviceTableComponents`: the block store is a thin
adapter over the slabs (so checkpoint / tablet iteration / host-protocol fallbacks keep working),
the Table API is served by kernels over the replicated pointer table, and the control protocol
gains four messages:

    TablePeerMap        driver -> executors : every owner's slab locations + IPC handles
    TableQuiesce/Resume driver -> executors : batch-granular access gate of ONE table (the device
                                              twin of the per-block ownership locks)
    DeviceMoveAlloc     driver -> receiver  : allocate the slabs of a LIST of blocks
    DeviceMoveSend      driver -> sender    : one bulk-TMA launch: old -> new slabs over NVLink + local snapshots
    DeviceOwnershipFlip driver -> executors : rewrite pointer-table entries (one launch) + fence; nobody is stopped
    DeviceMoveDrain     driver -> sender    : forward late updates (new += old - snapshot), free the old slabs

Reference parity: Tables.initTable (Tables.java:79-133), OwnershipCache.update
(OwnershipCache.java:195-244), MigrationExecutor (MigrationExecutor.java:134-414).
"""
from __future__ import annotations

import logging
import threading
from typing import Any, Iterator, List, Optional, Tuple

import numpy as np

from ..utils.concurrent import CompletedFuture, ReadWriteLock, ResultFuture
from .block_store import Block, BlockStore, Tablet
from .config import ExecutorConfiguration, TableConfiguration
from .executor import Executor
from .table import TableComponents, TableImpl
from .transport import Transport

LOG = logging.getLogger(__name__)


class DeviceContext:
    """Per-executor GPU state."""

    def __init__(self, device: int, executor_id: str, segment_bytes: Optional[int] = None) -> None:
        import torch
        from ..parallel.peer import PeerArena, PeerGroup
        self.device = int(device)
        # scoped: the creating thread (a launcher / test thread that also drives other GPUs) keeps its own device
        with torch.cuda.device(self.device):
            self.arena = PeerArena(self.device, PeerGroup(), segment_bytes=segment_bytes, self_key=executor_id)
            self.side_stream = torch.cuda.Stream(device=self.device)
        self.torch_device = torch.device("cuda", self.device)
        # GPU twin of the resource-typed task-unit slots (LocalTaskUnitScheduler.java:36-37: CPU x1, NET x2): the
        # jobs whose worker tasklets are active on THIS GPU split its SMs -- every job's kernels are launched with
        # grids sized for 1 / (active jobs) of the SMs (native thread-local `set_sm_share`), so a co-located job's
        # kernels run next to them instead of queueing behind a machine-filling grid.
        self._active_jobs: dict = {}
        self._jobs_lock = threading.Lock()

    def job_enter(self, job_id: str) -> None:
        with self._jobs_lock:
            self._active_jobs[job_id] = self._active_jobs.get(job_id, 0) + 1

    def job_leave(self, job_id: str) -> None:
        with self._jobs_lock:
            n = self._active_jobs.get(job_id, 0) - 1
            if n <= 0:
                self._active_jobs.pop(job_id, None)
            else:
                self._active_jobs[job_id] = n

    def sm_share(self, explicit: float = 0.0) -> float:
        """SM share of one job's kernels on this GPU: `-gpu_sm_share` if given, else 1 / (jobs active here)."""
        if explicit > 0:
            return min(1.0, explicit)
        with self._jobs_lock:
            return 1.0 / max(1, len(self._active_jobs))

    def apply_sm_share(self, explicit: float = 0.0) -> float:
        """Set the calling (tasklet) thread's launch budget; called at the start of every task unit."""
        from ..ops import native
        share = self.sm_share(explicit)
        native().set_sm_share(share)
        return share

    def tasklet_stream(self, priority: int = 0):
        """A fresh non-blocking stream for one tasklet: tasklets of different jobs (or of executors that share a
        GPU) overlap on the device instead of serialising on the legacy default stream -- the GPU half of the
        reference's multi-job co-scheduling (LocalTaskUnitScheduler.java:36-37 gives CPU x1 / NET x2 slots; here the
        hardware scheduler interleaves the streams' kernels)."""
        import torch
        # CUDA has two levels: 0 (default) and -1 (high); a job submitted with -job_priority high gets its kernels
        # scheduled ahead of co-located jobs' when both have work queued
        s = torch.cuda.Stream(device=self.device, priority=-1 if priority < 0 else 0)
        with self._jobs_lock:
            self._streams = getattr(self, "_streams", [])
            self._streams.append(s)              # strong reference until the tasklet releases it (release_stream)
        return s

    def release_stream(self, stream) -> None:
        with self._jobs_lock:
            self._streams = [x for x in getattr(self, "_streams", []) if x is not stream]

    def fence_streams(self) -> None:
        """Wait for the work that is ALREADY queued on this GPU's tasklet streams (+ the default and side streams) --
        and only for that.  A device-wide synchronize would also wait for what the (never stopped) workers keep
        enqueueing while it waits: with four busy jobs an ownership flip took 2.2 s instead of 27 ms."""
        import torch
        with self._jobs_lock:
            streams = list(getattr(self, "_streams", []))
        with torch.cuda.device(self.device):
            streams += [torch.cuda.default_stream(self.torch_device), self.side_stream]
            events = []
            for s in streams:
                ev = torch.cuda.Event()
                ev.record(s)
                events.append(ev)
            for ev in events:
                ev.synchronize()


class DeviceBlock(Block):
    """Block interface over one HBM slab (used by checkpoint, tablet iteration, tests)."""

    def __init__(self, block_id: int, dtable) -> None:
        self.block_id = block_id
        self._t = dtable

    def _keys(self) -> np.ndarray:
        return self._t.keys_of_block(self.block_id)

    def export_items(self):
        return self._t.export_block(self.block_id)

    def import_items(self, keys, values) -> None:
        self._t.import_block(self.block_id, np.asarray(keys), np.asarray(values))

    def items(self) -> Iterator[Tuple[Any, Any]]:
        k, v = self.export_items()
        return iter([(int(a), b) for a, b in zip(k, v)])

    def num_items(self) -> int:
        return len(self._keys())

    def multi_get_or_init(self, keys):
        import torch
        return self._t.pull(torch.as_tensor(np.asarray(keys), dtype=torch.int64, device=self._t.device)).cpu().numpy()

    multi_get = multi_get_or_init

    def get(self, key):
        return self.multi_get_or_init([key])[0]

    get_or_init = get

    def multi_put(self, keys, values):
        import torch
        self._t.put_rows(torch.as_tensor(np.asarray(keys), dtype=torch.int64, device=self._t.device),
                         torch.as_tensor(np.asarray(values), device=self._t.device))   # put_rows keeps int32 bit-exact

    def put(self, key, value):
        old = self.get(key)
        self.multi_put([key], [value])
        return old

    def multi_update(self, keys, deltas):
        import torch
        k = torch.as_tensor(np.asarray(keys), dtype=torch.int64, device=self._t.device)
        self._t.push(k, torch.as_tensor(np.asarray(deltas), device=self._t.device).float().contiguous())
        return self._t.pull(k).cpu().numpy()

    def update(self, key, delta):
        return self.multi_update([key], [delta])[0]


class DeviceBlockStore(BlockStore):
    """BlockStore facade over the HBM slabs of a device table (checkpoint, tablet iteration, tests see blocks; the
    data plane bypasses it)."""

    def __init__(self, dtable) -> None:
        super().__init__(lambda b: DeviceBlock(b, dtable))
        self._t = dtable

    def sync_with_table(self) -> None:
        with self._lock:
            local = set(self._t.local_block_ids())
            for b in list(self._blocks):
                if b not in local:
                    self._blocks.pop(b)
            for b in local:
                if b not in self._blocks:
                    self._blocks[b] = DeviceBlock(b, self._t)


class DeviceTableComponents(TableComponents):
    """Per-table state of a device executor: DeviceTable, ownership cache, read/write gate (device twin of the
    per-block ownership locks), ready event."""

    def __init__(self, conf: TableConfiguration, executor: "DeviceExecutor") -> None:
        super().__init__(conf, executor)
        self.dtable = None                    # created in init (needs the block owners)
        self.gate = ReadWriteLock()           # batch-granular access gate (quiesce for migration)
        self.ready = threading.Event()        # set once the peer map is complete
        self._executor = executor

    def create_device_table(self, block_owners: List[str]) -> None:
        from .device_table import DeviceTable
        ctx = self._executor.device_ctx
        self.dtable = DeviceTable(self.conf, ctx.arena, None, block_owners=block_owners, init=True,
                                  collective=False)
        self.block_store = DeviceBlockStore(self.dtable)
        self.block_store.sync_with_table()
        self.tablet = Tablet(self.block_store)

    def close(self) -> None:
        if self.dtable is not None:
            self.dtable.close()
            self.dtable = None


class DeviceTableImpl(TableImpl):
    """Table API served by kernels: multi-key ops are ONE launch whatever the owners are."""

    def __init__(self, comps: DeviceTableComponents, executor) -> None:
        super().__init__(comps, executor)
        self._dc = comps

    @property
    def device_table(self):
        self._dc.ready.wait()
        return self._dc.dtable

    def _keys(self, keys):
        import torch
        if isinstance(keys, torch.Tensor):
            return keys.to(self.device_table.device, dtype=torch.int64)
        return torch.as_tensor(np.asarray(keys, dtype=np.int64), device=self.device_table.device)

    def _vals(self, values):
        import torch
        if isinstance(values, torch.Tensor):
            return values.to(self.device_table.device).contiguous()
        dt = torch.float32 if self.conf.dtype == "float32" else torch.int32
        return torch.as_tensor(np.asarray(values), dtype=dt, device=self.device_table.device).contiguous()

    def _gated(self, fn):
        self._dc.gate.acquire_read()
        try:
            return fn()
        finally:
            self._dc.gate.release_read()

    def multi_get_or_init(self, keys) -> ResultFuture:
        t = self.device_table
        return CompletedFuture(self._gated(lambda: t.pull(self._keys(keys)).cpu().numpy()))

    multi_get = multi_get_or_init

    def get_or_init(self, key) -> ResultFuture:
        return CompletedFuture(self.multi_get_or_init([key]).get()[0])

    get = get_or_init

    def multi_update(self, keys, deltas) -> ResultFuture:
        t = self.device_table
        self._gated(lambda: t.push(self._keys(keys), self._vals(deltas)))
        return CompletedFuture(None)

    def multi_update_no_reply(self, keys, deltas) -> None:
        self.multi_update(keys, deltas)

    def update(self, key, delta) -> ResultFuture:
        self.multi_update([key], np.asarray(delta)[None, :])
        return self.get_or_init(key)

    def update_no_reply(self, key, delta) -> None:
        self.multi_update([key], np.asarray(delta)[None, :])

    def multi_put(self, keys, values) -> ResultFuture:
        t = self.device_table
        self._gated(lambda: t.put_rows(self._keys(keys), self._vals(values)))   # int32 tables: bit-exact path
        import torch
        torch.cuda.synchronize(t.device)
        return CompletedFuture(None)

    def put(self, key, value) -> ResultFuture:
        old = self.get_or_init(key).get()
        self.multi_put([key], np.asarray(value)[None, :])
        return CompletedFuture(old)

    def put_no_reply(self, key, value) -> None:
        self.multi_put([key], np.asarray(value)[None, :])

    # device handles for fused trainers
    def pull_device(self, keys):
        return self.device_table.pull(self._keys(keys))

    def push_device(self, keys, deltas) -> None:
        self.device_table.push(self._keys(keys), deltas)


class DeviceExecutor(Executor):
    """Executor whose dense `backend=device` tables live in HBM slabs of a CUDA-IPC exported arena; handles the
    device flavour of table init, peer map, quiesce / resume, migration receive and ownership flip."""

    def __init__(self, executor_id: str, transport: Transport, conf: Optional[ExecutorConfiguration] = None,
                 driver_id: str = "driver", device: Optional[int] = None) -> None:
        conf = conf or ExecutorConfiguration()
        dev = device if device is not None else (conf.resource.device or 0)
        self._device_index = dev
        super().__init__(executor_id, transport, conf, driver_id)
        self.device_ctx = DeviceContext(dev, executor_id)

    def make_table_components(self, conf: TableConfiguration) -> TableComponents:
        if conf.dense and conf.backend == "device":
            return DeviceTableComponents(conf, self)
        return TableComponents(conf, self)

    def make_table(self, comps: TableComponents) -> TableImpl:
        if isinstance(comps, DeviceTableComponents):
            return DeviceTableImpl(comps, self)
        return TableImpl(comps, self)

    def _cuda(self):
        import torch
        torch.cuda.set_device(self._device_index)
        return torch

    def _sync_device(self) -> None:
        """Control-plane handlers work on the default stream, tasklets on their own non-blocking streams
        (`DeviceContext.tasklet_stream`): a device-wide sync is the hand-over point between the two."""
        import torch
        torch.cuda.synchronize(self.device_ctx.torch_device)

    # ---- table control (device flavour) ----
    def _on_TableInit(self, m):
        conf = TableConfiguration.from_dict(m["table_conf"])
        if not (conf.dense and conf.backend == "device"):
            return super()._on_TableInit(m)
        self._cuda()
        comps = self.make_table_components(conf)
        comps.ownership.init(m["block_owners"])
        comps.create_device_table(m["block_owners"])
        comps.table = self.make_table(comps)
        with self.tables._lock:
            self.tables._tables[conf.table_id] = comps
        self._sync_device()          # slab init kernels done before anybody (tasklet streams, peers) sees the table
        self.send_to_driver({"type": "TableInitAck", "op_id": m["op_id"], "table_id": conf.table_id,
                             "device_info": comps.dtable.local_info()})

    def _on_TablePeerMap(self, m):
        torch = self._cuda()
        comps = self.tables.get_components(m["table_id"])
        comps.dtable.apply_infos(m["infos"])
        self._sync_device()
        if comps.dtable.ptrs_complete():
            comps.ready.set()
        self.send_to_driver({"type": "TablePeerMapAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_TableQuiesce(self, m):
        torch = self._cuda()
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None and isinstance(comps, DeviceTableComponents):
            comps.gate.acquire_write()
            torch.cuda.synchronize(self.device_ctx.torch_device)
        self.send_to_driver({"type": "TableQuiesceAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_TableResume(self, m):
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None and isinstance(comps, DeviceTableComponents):
            comps.gate.release_write()
        self.send_to_driver({"type": "TableResumeAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    # ---- live device migration: alloc (receiver) / send (sender) / flip (everyone) / drain (sender) ----
    def _on_DeviceMoveAlloc(self, m):
        self._cuda()
        comps = self.tables.get_components(m["table_id"])
        locs = comps.dtable.alloc_recv(m["block_ids"])
        self.send_to_driver({"type": "DeviceMoveAllocAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "locs": locs})

    def _on_DeviceMoveAbort(self, m):
        self._cuda()
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None:
            comps.dtable.abort_recv(m["block_ids"])
        self.send_to_driver({"type": "DeviceMoveAbortAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_DeviceMoveSend(self, m):
        self._cuda()
        comps = self.tables.get_components(m["table_id"])
        locs = {int(b): tuple(loc) for b, loc in m["locs"].items()}
        ms = comps.dtable.send_blocks(locs, self.device_ctx.side_stream)
        self.send_to_driver({"type": "DeviceMoveSendAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "copy_ms": ms, "bytes": len(locs) * comps.dtable.block_bytes})

    def _on_DeviceOwnershipFlip(self, m):
        self._cuda()
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None:
            locs = {int(b): tuple(loc) for b, loc in m["locs"].items()}
            # one launch + a fence over the work already queued on this GPU's streams; nobody is stopped
            comps.dtable.flip_blocks(locs, m["new_owner"], fence=self.device_ctx.fence_streams)
            for b in locs:
                old = comps.ownership.owners()[b]
                comps.ownership._owners[b] = m["new_owner"]
                with comps.ownership._meta_lock:
                    comps.ownership._num_blocks[old] = comps.ownership._num_blocks.get(old, 1) - 1
                    comps.ownership._num_blocks[m["new_owner"]] = comps.ownership._num_blocks.get(m["new_owner"], 0) + 1
            comps.block_store.sync_with_table()
        self.send_to_driver({"type": "DeviceOwnershipFlipAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_DeviceMoveDrain(self, m):
        self._cuda()
        comps = self.tables.get_components(m["table_id"])
        n = comps.dtable.drain_blocks(m["block_ids"], self.device_ctx.side_stream)
        comps.block_store.sync_with_table()
        self.send_to_driver({"type": "DeviceMoveDrainAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "forwarded": n})

    def close(self) -> None:
        super().close()
