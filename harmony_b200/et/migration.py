"""Executor-side halves of block migration (host backend).

Protocol (reference: evaluator/impl/MigrationExecutor.java:107-132 MoveInit, :134-198 ownership
first hand-off for MUTABLE tables, :202-253 receiver assembling chunks, :255-295 acks,
:338-414 chunked sender; constants :51-54 = <=4 blocks in flight, chunk = `chunk_size` items):

  mutable   : sender --Ownership--> receiver (update cache, ARM latch) --OwnershipAck--> sender
              (update cache; report OwnershipMoved to the driver) ; sender streams Data chunks ;
              receiver putBlock + release latch --DataAck--> sender (remove block; report DataMoved)
  immutable : data and ownership travel together (Data chunks first, then both caches flip).

On the device backend the same state machine is driven by DeviceTable.migrate_block: the payload
is a P2P bulk-copy kernel on a side stream instead of Data messages.
"""
from __future__ import annotations

import logging
import threading
from typing import Any, Dict, Tuple

LOG = logging.getLogger(__name__)
MAX_CONCURRENT_MIGRATIONS = 4


def _chunks(keys, values, chunk_size: int):
    n = len(keys)
    if n == 0:
        yield keys, values
        return
    for i in range(0, n, chunk_size):
        yield keys[i:i + chunk_size], values[i:i + chunk_size]


class MigrationExecutor:
    """Executor halves of block migration: sender (<= 4 blocks in flight, ownership-first for mutable tables,
    chunked data messages) and receiver (install block, update cache, acks).  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/MigrationExecutor.java:107-132,134-198,202-253,255-295,338-414."""

    def __init__(self, executor) -> None:
        self._ex = executor
        self._sem = threading.Semaphore(MAX_CONCURRENT_MIGRATIONS)
        self._incoming: Dict[Tuple[str, int], Dict[str, Any]] = {}
        self._lock = threading.Lock()
        # dedicated sender thread (the reference uses 2): pool threads must never block on the
        # in-flight limit, the acks that release it are served by the same pool
        import queue
        self._queue: "queue.Queue" = queue.Queue()
        self._sender = threading.Thread(target=self._sender_loop, name=f"{executor.executor_id}-mig-sender",
                                        daemon=True)
        self._sender.start()

    # ---------------- sender side ----------------
    def on_move_init(self, m: Dict[str, Any]) -> None:
        for block_id in m["block_ids"]:
            self._queue.put((m["op_id"], m["table_id"], block_id, m["receiver_id"]))

    def _sender_loop(self) -> None:
        while True:
            item = self._queue.get()
            if item is None:
                return
            self._sem.acquire()              # <= MAX_CONCURRENT_MIGRATIONS blocks in flight
            self._move_one(*item)

    def _move_one(self, op_id: int, table_id: str, block_id: int, receiver: str) -> None:
        try:
            comps = self._ex.tables.get_components(table_id)
            if comps.conf.is_mutable:
                # ownership first
                self._ex.transport.send(receiver, {"type": "MigOwnership", "op_id": op_id, "table_id": table_id,
                                                   "block_id": block_id, "old": self._ex.executor_id,
                                                   "new": receiver})
            else:
                self._send_data(op_id, table_id, block_id, receiver, move_ownership_together=True)
        except Exception:  # noqa: BLE001
            LOG.exception("migration of block %s failed", block_id)
            self._sem.release()

    def on_ownership_ack(self, m: Dict[str, Any]) -> None:
        table_id, block_id = m["table_id"], m["block_id"]
        comps = self._ex.tables.get_components(table_id)
        comps.ownership.update(block_id, m["old"], m["new"])   # waits for in-flight local ops (write lock)
        self._ex.send_to_driver({"type": "OwnershipMoved", "op_id": m["op_id"], "table_id": table_id,
                                 "block_id": block_id})
        self._send_data(m["op_id"], table_id, block_id, m["new"], move_ownership_together=False)

    def _send_data(self, op_id: int, table_id: str, block_id: int, receiver: str,
                   move_ownership_together: bool) -> None:
        comps = self._ex.tables.get_components(table_id)
        block = comps.block_store.get(block_id)
        keys, values = block.export_items()
        total = len(keys)
        for k, v in _chunks(keys, values, max(1, comps.conf.chunk_size)):
            self._ex.transport.send(receiver, {"type": "MigData", "op_id": op_id, "table_id": table_id,
                                               "block_id": block_id, "num_total_items": total, "keys": k,
                                               "values": v, "sender_id": self._ex.executor_id,
                                               "receiver_id": receiver,
                                               "ownership_together": move_ownership_together})

    def on_data_ack(self, m: Dict[str, Any]) -> None:
        table_id, block_id = m["table_id"], m["block_id"]
        comps = self._ex.tables.get_components(table_id)
        try:
            if m["ownership_together"]:
                comps.ownership.update(block_id, self._ex.executor_id, m["receiver_id"])
            comps.block_store.remove_block(block_id)
            self._ex.send_to_driver({"type": "DataMoved", "op_id": m["op_id"], "table_id": table_id,
                                     "block_id": block_id, "ownership_together": m["ownership_together"]})
        finally:
            self._sem.release()

    # ---------------- receiver side ----------------
    def on_ownership(self, m: Dict[str, Any]) -> None:
        comps = self._ex.tables.get_components(m["table_id"])
        comps.ownership.update(m["block_id"], m["old"], m["new"])   # arms the incoming-block latch
        self._ex.transport.send(m["old"], {"type": "MigOwnershipAck", "op_id": m["op_id"],
                                           "table_id": m["table_id"], "block_id": m["block_id"],
                                           "old": m["old"], "new": m["new"]})

    def on_data(self, m: Dict[str, Any]) -> None:
        key = (m["table_id"], m["block_id"])
        with self._lock:
            st = self._incoming.setdefault(key, {"keys": [], "values": [], "n": 0})
            st["keys"].append(m["keys"])
            st["values"].append(m["values"])
            st["n"] += len(m["keys"])
            done = st["n"] >= m["num_total_items"]
            if done:
                self._incoming.pop(key)
        if not done:
            return
        comps = self._ex.tables.get_components(m["table_id"])
        block = comps.block_store.create_empty_block(m["block_id"])
        for k, v in zip(st["keys"], st["values"]):
            block.import_items(k, v)
        if m["ownership_together"]:
            comps.ownership.update(m["block_id"], m["sender_id"], self._ex.executor_id)
        if comps.ownership.is_incoming(m["block_id"]):
            comps.ownership.allow_access_to_block(m["block_id"])
        self._ex.transport.send(m["sender_id"], {"type": "MigDataAck", "op_id": m["op_id"],
                                                 "table_id": m["table_id"], "block_id": m["block_id"],
                                                 "receiver_id": self._ex.executor_id,
                                                 "ownership_together": m["ownership_together"]})
