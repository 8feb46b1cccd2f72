"""Centralized communication: named driver <-> executor byte/object channels multiplexed over the
control transport.

Reference: common/.../centcomm/CentCommConf.java:46-145 (per-client handler registration),
master/MasterSideCentCommMsgSender.java:64-74, slave/SlaveSideCentCommMsgSender.java:64-78,
ns/CentCommMsgHandler.java:40-86 (route by client class name), common/src/main/avro/centcomm.avsc
(`CentCommMsg{sourceId, clientClassName, data}`).
"""
from __future__ import annotations

from typing import Any, Callable

Handler = Callable[[str, Any], None]     # (source id, data)


class MasterSideCentComm:
    """Driver side of the named message channels between the driver and its executors.  Reference:
    common/src/main/java/edu/snu/cay/common/centcomm/master/MasterSideCentCommMsgSender.java,
    CentCommConf.java:46-145."""

    def __init__(self, et_master) -> None:
        self._m = et_master

    def register_handler(self, client_name: str, handler: Handler) -> None:
        self._m.centcomm_handlers[client_name] = lambda m: handler(m["source_id"], m["data"])

    def send(self, client_name: str, executor_id: str, data: Any) -> None:
        self._m.send(executor_id, {"type": "CentComm", "client": client_name, "source_id": self._m.driver_id,
                                   "data": data})


class SlaveSideCentComm:
    """Executor side of CentComm.  Reference:
    common/src/main/java/edu/snu/cay/common/centcomm/slave/SlaveSideCentCommMsgSender.java."""

    def __init__(self, executor) -> None:
        self._ex = executor

    def register_handler(self, client_name: str, handler: Handler) -> None:
        self._ex.user_services[client_name] = lambda m: handler(m["source_id"], m["data"])

    def send(self, client_name: str, data: Any) -> None:
        self._ex.send_to_driver({"type": "CentComm", "client": client_name, "source_id": self._ex.executor_id,
                                 "data": data})
