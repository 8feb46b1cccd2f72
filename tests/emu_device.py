"""Host-memory stand-in for `DeviceExecutor`, for CPU tests of the DRIVER side of the device-table protocol.

It speaks exactly the control messages of the HBM backend -- `device_info` in TableInitAck, TablePeerMap,
TableQuiesce / TableResume, DeviceMoveAlloc / Send / Drain, DeviceOwnershipFlip (et/device_executor.py) -- but keeps the rows in the
host dense blocks: "peer memory" is a process-wide registry of executors; the live move keeps a snapshot at the
sender, fences each moved block through its ownership write lock and forwards late updates at the drain.  The production GPU code is not involved; what is exercised is AllocatedTable's device paths
(`_exchange_peer_map`, `_device_move` incl. its abort path, late subscribers) and everything above them.
"""
from typing import Any, Callable, Dict, List

import numpy as np

from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.executor import Executor


class EmuDeviceExecutor(Executor):
    registry: Dict[str, "EmuDeviceExecutor"] = {}
    fail_recv_on: set = set()           # executor ids whose DeviceMoveAlloc raises (fault injection)
    fail_send_on: set = set()           # executor ids whose DeviceMoveSend raises
    aborted: list = []                  # (executor id, block ids) of every DeviceMoveAbort received

    def __init__(self, executor_id, transport, conf=None, driver_id="driver", **_: Any) -> None:
        super().__init__(executor_id, transport, conf, driver_id)
        EmuDeviceExecutor.registry[executor_id] = self
        self._gates: Dict[str, List[Callable[[], None]]] = {}
        self.peer_maps: Dict[str, list] = {}

    @staticmethod
    def _is_dev(conf: TableConfiguration) -> bool:
        return bool(conf.dense and conf.backend == "device")

    def _on_TableInit(self, m):
        conf = TableConfiguration.from_dict(m["table_conf"])
        if not self._is_dev(conf):
            return super()._on_TableInit(m)
        self.tables.init_table(conf, m["block_owners"])
        mine = [b for b, o in enumerate(m["block_owners"]) if o == self.executor_id]
        info = {"key": self.executor_id, "blocks": {b: (self.executor_id, 0, b) for b in mine}, "handles": {},
                "ptr_tbl": None}
        self.send_to_driver({"type": "TableInitAck", "op_id": m["op_id"], "table_id": conf.table_id,
                             "device_info": info})

    def _on_TablePeerMap(self, m):
        self.peer_maps[m["table_id"]] = m["infos"]
        self.send_to_driver({"type": "TablePeerMapAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_TableQuiesce(self, m):
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None and self._is_dev(comps.conf):
            # the table gate: exclusive on every block, i.e. no accessor is inside the table
            self._gates[m["table_id"]] = [comps.ownership.hold_write_lock(b)
                                          for b in range(comps.conf.num_total_blocks)]
        self.send_to_driver({"type": "TableQuiesceAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_TableResume(self, m):
        for unlock in self._gates.pop(m["table_id"], []):
            unlock()
        self.send_to_driver({"type": "TableResumeAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    # ---- live migration: alloc / send (+ snapshot) / flip (+ per-block fence) / drain ----
    def _on_DeviceMoveAlloc(self, m):
        if self.executor_id in EmuDeviceExecutor.fail_recv_on:
            raise MemoryError("emulated: receiver arena is full")
        locs = {b: (self.executor_id, 0, b, {}) for b in m["block_ids"]}
        self.send_to_driver({"type": "DeviceMoveAllocAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "locs": locs})

    def _on_DeviceMoveAbort(self, m):
        EmuDeviceExecutor.aborted.append((self.executor_id, sorted(int(b) for b in m["block_ids"])))
        self.send_to_driver({"type": "DeviceMoveAbortAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_DeviceMoveSend(self, m):
        if self.executor_id in EmuDeviceExecutor.fail_send_on:
            raise RuntimeError("emulated: bulk copy timed out")
        comps = self.tables.get_components(m["table_id"])
        self._snap = getattr(self, "_snap", {})
        for b, loc in m["locs"].items():
            b = int(b)
            keys, vals = comps.block_store.get(b).export_items()          # the "bulk copy" + the local snapshot
            dst = EmuDeviceExecutor.registry[loc[0]].tables.get_components(m["table_id"])
            dst.block_store.put_block(b, keys, [np.array(v, copy=True) for v in vals])
            self._snap[(m["table_id"], b)] = {k: np.array(v, copy=True) for k, v in zip(keys, vals)}
        self.send_to_driver({"type": "DeviceMoveSendAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "copy_ms": 0.0, "bytes": 0})

    def _on_DeviceOwnershipFlip(self, m):
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None:
            for b in m["locs"]:
                b = int(b)
                # the fence of THIS block: in-flight accessors hold its read lock; nothing else is stopped
                unlock = comps.ownership.hold_write_lock(b)
                try:
                    old = comps.ownership.owners()[b]
                    comps.ownership._owners[b] = m["new_owner"]
                    with comps.ownership._meta_lock:
                        nb = comps.ownership._num_blocks
                        nb[old] = nb.get(old, 1) - 1
                        nb[m["new_owner"]] = nb.get(m["new_owner"], 0) + 1
                finally:
                    unlock()
        self.send_to_driver({"type": "DeviceOwnershipFlipAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def _on_DeviceMoveDrain(self, m):
        comps = self.tables.get_components(m["table_id"])
        forwarded = 0
        for b in m["block_ids"]:
            b = int(b)
            snap = self._snap.pop((m["table_id"], b))
            keys, vals = comps.block_store.get(b).export_items()
            owner = comps.ownership.owners()[b]
            dst = EmuDeviceExecutor.registry[owner].tables.get_components(m["table_id"])
            late = [(k, np.asarray(v) - snap[k]) for k, v in zip(keys, vals) if k in snap and np.any(np.asarray(v) != snap[k])]
            late += [(k, np.asarray(v)) for k, v in zip(keys, vals) if k not in snap]
            if late:                                                      # new += old - snapshot
                dst.block_store.get(b).multi_update([k for k, _ in late], [d for _, d in late])
                forwarded += len(late)
            comps.block_store.remove_block(b)
        self.send_to_driver({"type": "DeviceMoveDrainAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "forwarded": forwarded})

    def shutdown(self):
        EmuDeviceExecutor.registry.pop(self.executor_id, None)
        return super().shutdown()
