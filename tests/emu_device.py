"""Host-memory stand-in for `DeviceExecutor`, for CPU tests of the DRIVER side of the device-table protocol.

It speaks exactly the control messages of the HBM backend -- `device_info` in TableInitAck, TablePeerMap,
TableQuiesce / TableResume, DeviceMoveRecv, DeviceOwnershipFlip (et/device_executor.py) -- but keeps the rows in the
host dense blocks: "peer memory" is a process-wide registry of executors, the table gate is the write lock of every
block's ownership lock.  The production GPU code is not involved; what is exercised is AllocatedTable's device paths
(`_exchange_peer_map`, `_device_move` incl. its abort path, late subscribers) and everything above them.
"""
from typing import Any, Callable, Dict, List

from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.executor import Executor


class EmuDeviceExecutor(Executor):
    registry: Dict[str, "EmuDeviceExecutor"] = {}
    fail_recv_on: set = set()           # executor ids whose DeviceMoveRecv raises (fault injection)

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

    def _on_DeviceMoveRecv(self, m):
        if self.executor_id in EmuDeviceExecutor.fail_recv_on:
            raise MemoryError("emulated: receiver arena is full")
        comps = self.tables.get_components(m["table_id"])
        locs = {}
        for b in m["block_ids"]:
            owner = comps.ownership.owners()[b]
            src = EmuDeviceExecutor.registry[owner].tables.get_components(m["table_id"])
            keys, vals = src.block_store.get(b).export_items()          # the "P2P bulk copy"
            comps.block_store.put_block(b, keys, vals)
            locs[b] = (self.executor_id, 0, b, {})
        self.send_to_driver({"type": "DeviceMoveRecvAck", "op_id": m["op_id"], "table_id": m["table_id"],
                             "locs": locs})

    def _on_DeviceOwnershipFlip(self, m):
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None:
            for b in m["locs"]:
                b = int(b)
                old = comps.ownership.owners()[b]
                # the gate already excludes accessors: update the cache the way DeviceExecutor does (no latch)
                comps.ownership._owners[b] = m["new_owner"]
                with comps.ownership._meta_lock:
                    nb = comps.ownership._num_blocks
                    nb[old] = nb.get(old, 1) - 1
                    nb[m["new_owner"]] = nb.get(m["new_owner"], 0) + 1
                if old == self.executor_id and m["new_owner"] != self.executor_id:
                    comps.block_store.remove_block(b)
        self.send_to_driver({"type": "DeviceOwnershipFlipAck", "op_id": m["op_id"], "table_id": m["table_id"]})

    def shutdown(self):
        EmuDeviceExecutor.registry.pop(self.executor_id, None)
        return super().shutdown()
