"""An ET executor: one per GPU (or per CPU worker in the host backend).

Wires the evaluator-side components of the reference (services/et/.../evaluator/impl/*):
Tables, OwnershipCache (inside TableComponents), RemoteAccessOpSender/Handler + CommManager,
MigrationExecutor, ChkpManagerSlave, TaskletRuntime, LocalTaskUnitScheduler, MetricCollector and
the message demultiplexer (MessageHandlerImpl.java:106-383: 7 pools by message kind; here one
dispatcher thread per endpoint hands work to a shared pool, table access ops go to the per-block
FIFO comm threads).
"""
from __future__ import annotations

import logging
import threading
from typing import Any, Dict, Optional

from ..utils.concurrent import new_fixed_thread_pool
from .checkpoint import ChkpManagerSlave
from .config import ExecutorConfiguration, TableConfiguration, TaskletConfiguration
from .metric import MetricCollector
from .migration import MigrationExecutor
from .remote_access import RemoteAccessOpHandler, RemoteAccessOpSender
from .table import TableAccessor, TableComponents, TableImpl, Tables
from .tasklet import LocalTaskUnitScheduler, TaskletRuntime
from .transport import NetworkError, Transport

LOG = logging.getLogger(__name__)


class Executor:
    """One ET executor: table registry, remote-access handler / sender, migration executor, checkpoint slave,
    tasklet runtime, task-unit scheduler, metric collector, message demux.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/MessageHandlerImpl.java:55-69,106-138,194-289,
    Tables.java:79-133."""

    def __init__(self, executor_id: str, transport: Transport, conf: Optional[ExecutorConfiguration] = None,
                 driver_id: str = "driver") -> None:
        self.executor_id = executor_id
        self.driver_id = driver_id
        self.transport = transport
        self.conf = conf or ExecutorConfiguration()
        self.pool = new_fixed_thread_pool(16, f"{executor_id}-pool")
        self.tables = Tables(self)
        self.table_accessor = TableAccessor(self.tables)
        self.sender = RemoteAccessOpSender(self)
        self.handler = RemoteAccessOpHandler(self, self.conf.remote_access.num_handler_threads,
                                             self.conf.remote_access.handler_queue_size)
        self.migration = MigrationExecutor(self)
        self.chkp = ChkpManagerSlave(self, self.conf.chkp_temp_path, self.conf.chkp_commit_path)
        self.task_unit_scheduler = LocalTaskUnitScheduler(self)
        self.metric_collector = MetricCollector(self)
        self.tasklet_runtime = TaskletRuntime(self, self.conf.num_tasklets)
        self.user_services: Dict[str, Any] = {}       # CentComm-style named services
        self._closed = threading.Event()
        self.device_ctx = None                         # set by the GPU launcher (arena, peer group)
        transport.register(executor_id, self.on_message)

    # ---- factories (overridden by the device executor) ----
    def make_table_components(self, conf: TableConfiguration) -> TableComponents:
        return TableComponents(conf, self)

    def make_table(self, comps: TableComponents) -> TableImpl:
        return TableImpl(comps, self)

    # ---- outbound helpers ----
    def send_to_driver(self, msg: Dict[str, Any]) -> None:
        msg.setdefault("src", self.executor_id)
        try:
            self.transport.send(self.driver_id, msg)
        except NetworkError:
            LOG.warning("[%s] driver unreachable for %s", self.executor_id, msg.get("type"))

    def send_ownership_sync_ack(self, op_id: int, table_id: str, deleted_executor_id: str) -> None:
        self.send_to_driver({"type": "OwnershipSyncAck", "op_id": op_id, "table_id": table_id,
                             "deleted_executor_id": deleted_executor_id})

    # ---- inbound demux ----
    def on_message(self, m: Dict[str, Any]) -> None:
        t = m["type"]
        if t == "TableAccessReq":
            self.handler.on_table_access_req(m)
        elif t == "TableAccessRes":
            self.sender.on_table_access_res(m)
        elif t == "TableAccessReqBatch":
            self.handler.on_table_access_req_batch(m)
        elif t == "TableAccessResBatch":
            self.sender.on_table_access_res_batch(m)
        elif t == "TaskUnitReady":
            self.task_unit_scheduler.on_ready(m)
        elif t == "TaskletCustom" and m.get("inline"):
            self._on_TaskletCustom(m)            # non-blocking by the sender's contract (barrier releases)
        else:
            fn = getattr(self, "_on_" + t, None)
            if fn is None:
                svc = self.user_services.get(m.get("client"))
                if t == "CentComm" and svc is not None:
                    self.pool.submit(svc, m)
                    return
                LOG.warning("[%s] unhandled message type %s", self.executor_id, t)
                return
            self.pool.submit(self._guarded, fn, m)

    # control messages the driver fans out and waits on: a failing handler must fail the driver-side future
    _CONTROL_OPS = {"TableInit", "TableLoad", "TableDrop", "OwnershipSync", "TablePeerMap", "TableQuiesce",
                    "TableResume", "DeviceMoveAlloc", "DeviceMoveSend", "DeviceMoveDrain", "DeviceMoveAbort",
                    "DeviceOwnershipFlip"}

    def _guarded(self, fn, m: Dict[str, Any]) -> None:
        try:
            fn(m)
        except Exception as e:  # noqa: BLE001
            if m.get("type") in self._CONTROL_OPS and "op_id" in m:
                LOG.exception("[%s] control op %s failed", self.executor_id, m.get("type"))
                self.send_to_driver({"type": "ControlOpFailed", "op_id": m["op_id"], "orig_type": m["type"],
                                     "table_id": m.get("table_id"), "error": f"{type(e).__name__}: {e}"})
            else:
                raise

    # -- table control --
    def _on_TableInit(self, m):
        conf = TableConfiguration.from_dict(m["table_conf"])
        self.tables.init_table(conf, m["block_owners"])
        self.send_to_driver({"type": "TableInitAck", "op_id": m["op_id"], "table_id": conf.table_id})

    def _on_TableLoad(self, m):
        from ..common.dataloader import run_bulk_load
        n = run_bulk_load(self, m["table_id"], m["file_splits"])
        self.send_to_driver({"type": "TableLoadAck", "op_id": m["op_id"], "table_id": m["table_id"], "num_items": n})

    def _on_TableDrop(self, m):
        table_id = m["table_id"]
        self.sender.wait_ops_to_be_flushed(table_id)
        self.handler.wait_ops_to_be_flushed(table_id)
        comps = self.tables.try_get_components(table_id)
        if comps is not None:
            comps.ownership.complete_all_ongoing_sync()
            self.tables.remove(table_id)
        self.handler.forget_table(table_id)
        self.send_to_driver({"type": "TableDropAck", "op_id": m["op_id"], "table_id": table_id})

    def _on_OwnershipUpdate(self, m):
        comps = self.tables.try_get_components(m["table_id"])
        if comps is not None:
            comps.ownership.update(m["block_id"], m["old"], m["new"])

    def _on_OwnershipSync(self, m):
        comps = self.tables.try_get_components(m["table_id"])
        if comps is None:
            self.send_ownership_sync_ack(m["op_id"], m["table_id"], m["deleted_executor_id"])
        else:
            comps.ownership.sync_unassociation(m["op_id"], m["deleted_executor_id"])

    # -- checkpoint --
    def _on_ChkpStart(self, m):
        try:
            blocks = self.chkp.checkpoint(m["chkp_id"], m["table_id"], m.get("sampling_ratio", 1.0))
        except Exception as e:  # noqa: BLE001  (full disk, dropped table ...): fail the driver-side future, loudly
            LOG.exception("[%s] checkpoint %s failed", self.executor_id, m["chkp_id"])
            self.send_to_driver({"type": "ChkpFailed", "chkp_id": m["chkp_id"], "executor_id": self.executor_id,
                                 "error": f"{type(e).__name__}: {e}"})
            return
        self.send_to_driver({"type": "ChkpDone", "chkp_id": m["chkp_id"], "executor_id": self.executor_id,
                             "block_ids": blocks})

    def _on_ChkpCommit(self, m):
        self.chkp.commit(m["chkp_id"])
        self.send_to_driver({"type": "ChkpCommitAck", "chkp_id": m["chkp_id"], "executor_id": self.executor_id})

    def _on_ChkpLoad(self, m):
        try:
            n = self.chkp.load(m["chkp_id"], m["table_id"], m["block_ids"], m.get("committed", False))
        except Exception as e:  # noqa: BLE001  (missing block file, multi_put timeout ...)
            LOG.exception("[%s] restore of %s failed", self.executor_id, m["chkp_id"])
            self.send_to_driver({"type": "ChkpFailed", "chkp_id": m["chkp_id"], "op_id": m["op_id"],
                                 "executor_id": self.executor_id, "error": f"{type(e).__name__}: {e}"})
            return
        self.send_to_driver({"type": "ChkpLoadDone", "chkp_id": m["chkp_id"], "op_id": m["op_id"],
                             "executor_id": self.executor_id, "num_items": n})

    # -- migration --
    def _on_MoveInit(self, m):
        self.migration.on_move_init(m)

    def _on_MigOwnership(self, m):
        self.migration.on_ownership(m)

    def _on_MigOwnershipAck(self, m):
        self.migration.on_ownership_ack(m)

    def _on_MigData(self, m):
        self.migration.on_data(m)

    def _on_MigDataAck(self, m):
        self.migration.on_data_ack(m)

    # -- metric / tasklet --
    def _on_MetricControl(self, m):
        if m["command"] == "Start":
            self.metric_collector.start(m.get("flush_period_ms", 0))
        else:
            self.metric_collector.stop()

    def _on_TaskletStart(self, m):
        try:
            self.tasklet_runtime.start_tasklet(TaskletConfiguration.from_dict(m["conf"]))
        except Exception as e:  # noqa: BLE001
            LOG.exception("cannot start tasklet")
            self.send_to_driver({"type": "TaskletStatus", "tasklet_id": m["conf"]["tasklet_id"],
                                 "executor_id": self.executor_id, "status": "Failed", "result": None,
                                 "error": f"{type(e).__name__}: {e}"})

    def _on_TaskletStop(self, m):
        self.tasklet_runtime.stop_tasklet(m["tasklet_id"])

    def _on_TaskletCustom(self, m):
        self.tasklet_runtime.on_custom_msg(m["tasklet_id"], m["payload"])

    def _on_ExecutorClose(self, m):
        self.close()
        try:
            self.transport.send(self.driver_id, {"type": "ExecutorClosed", "executor_id": self.executor_id,
                                                 "src": self.executor_id})
        except NetworkError:
            pass

    # ---- lifecycle ----
    def close(self) -> None:
        """ContextStopHandler: close tasklets and COMMIT all local checkpoints."""
        if self._closed.is_set():
            return
        self._closed.set()
        self.tasklet_runtime.close()
        try:
            self.chkp.commit_all_local_chkps()
        except Exception:  # noqa: BLE001
            LOG.exception("commit on close failed")
        self.handler.close()

    def shutdown(self) -> None:
        self.close()
        self.transport.unregister(self.executor_id)
        self.pool.shutdown(wait=False)

    def wait_closed(self, timeout: Optional[float] = None) -> bool:
        return self._closed.wait(timeout)
