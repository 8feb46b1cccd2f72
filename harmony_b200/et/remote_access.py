"""Remote table access on the host backend: op descriptors, per-block FIFO comm threads,
the sender (origin side) and the handler (owner side).

Reference behaviour:
  * CommManager.java:67-154      -- N threads, op -> thread by blockId % N  => per-block FIFO
  * RemoteAccessOpSender.java    -- register op, encode+send, retry on NetworkException after
                                    asking the driver for the owner (:124-163), complete futures
                                    on the response (:351-392), waitOpsTobeFlushed (:457-476)
  * RemoteAccessOpHandler.java   -- execute under the ownership read lock (:236-321), redirect
                                    to the real owner when stale (:225-231), redirect to the
                                    driver FallbackManager when the table is unknown / flushing
                                    (:119-129), op tracker PROCESSING -> FLUSHING -> FLUSHED (:446-555)
On GPUs none of this runs for pull/push: the kernels in ops/csrc/table_ops.cu resolve owners
from the HBM pointer table and touch the owner's memory directly.
"""
from __future__ import annotations

import itertools
import logging
import queue
import threading
import time
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from ..utils.concurrent import ResultFuture
from ..utils.state_machine import StateMachine
from .exceptions import DataAccessFailedError
from .transport import NetworkError

LOG = logging.getLogger(__name__)

PUT, PUT_IF_ABSENT, GET, GET_OR_INIT, REMOVE, UPDATE = "PUT", "PUT_IF_ABSENT", "GET", "GET_OR_INIT", "REMOVE", "UPDATE"
RESEND_INTERVAL_S = 0.1   # RemoteAccessOpSender.java:47


class RemoteAccessOpStat:
    """GET-response byte / request counters fed to the metric service.

    (The reference never increments the request counter -- SURVEY.md Appendix E.1; we do.)"""

    def __init__(self) -> None:
        self._lock = threading.Lock()
        self._bytes = 0
        self._count = 0

    def inc_count_sent_get_req(self, n: int = 1) -> None:
        with self._lock:
            self._count += n

    def inc_bytes_received_get_resp(self, n: int) -> None:
        with self._lock:
            self._bytes += n

    def get_and_reset(self):
        with self._lock:
            out = (self._count, self._bytes)
            self._count = self._bytes = 0
            return out


class SingleKeyDataOpResult(ResultFuture):
    """Future of a single-key op.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/SingleKeyDataOpResult.java."""

    pass


class MultiKeyDataOpResult(ResultFuture):
    """Aggregates per-block partial results back into the caller's key order."""

    def __init__(self, n_keys: int, n_blocks: int, dense: bool) -> None:
        super().__init__()
        self._remaining = n_blocks
        self._dense = dense
        self._values: List[Any] = [None] * n_keys
        self._ok = True
        self._agg_lock = threading.Lock()
        if n_blocks == 0:
            self.on_completed([] if not dense else np.zeros((0, 0), dtype=np.float32))

    def on_block_completed(self, positions: Sequence[int], values: Optional[Sequence[Any]], success: bool) -> None:
        with self._agg_lock:
            if values is not None:
                for p, v in zip(positions, values):
                    self._values[p] = v
            self._ok = self._ok and success
            self._remaining -= 1
            done = self._remaining == 0
        if done:
            if not self._ok:
                self.on_failed(DataAccessFailedError("multi-key op failed on at least one block"))
            elif self._dense and self._values and all(v is not None for v in self._values):
                self.on_completed(np.stack(self._values))
            else:
                self.on_completed(self._values)


class RemoteDataOp:
    """One (table, block) operation travelling between the origin and the owner."""
    __slots__ = ("op_id", "orig_id", "table_id", "op_type", "block_id", "keys", "values", "reply_required",
                 "single", "callback", "is_local", "batch")

    def __init__(self, op_id, orig_id, table_id, op_type, block_id, keys, values, reply_required, single,
                 callback=None, is_local=False) -> None:
        self.op_id, self.orig_id, self.table_id, self.op_type = op_id, orig_id, table_id, op_type
        self.block_id, self.keys, self.values = block_id, keys, values
        self.reply_required, self.single, self.callback, self.is_local = reply_required, single, callback, is_local
        self.batch: Optional["_ReplyBatch"] = None      # owner side: replies of one request batch travel together

    def to_msg(self) -> Dict[str, Any]:
        return {"type": "TableAccessReq", "op_id": self.op_id, "orig_id": self.orig_id, "table_id": self.table_id,
                "op_type": self.op_type, "block_id": self.block_id, "keys": self.keys, "values": self.values,
                "reply_required": self.reply_required, "single": self.single}

    @staticmethod
    def from_msg(m: Dict[str, Any]) -> "RemoteDataOp":
        return RemoteDataOp(m["op_id"], m["orig_id"], m["table_id"], m["op_type"], m["block_id"], m["keys"],
                            m.get("values"), m["reply_required"], m["single"])


class _ReplyBatch:
    """Owner side of a `TableAccessReqBatch`: the block ops of one multi-key request that target this executor were
    sent in ONE message; their results go back in one `TableAccessResBatch` once the last of them has executed
    (ops that had to be forwarded to a new owner leave the batch and are answered individually)."""

    def __init__(self, expected: int, orig_id: str, executor) -> None:
        self._expected, self._orig, self._ex = expected, orig_id, executor
        self._results: List[Dict[str, Any]] = []
        self._lock = threading.Lock()

    def _maybe_send(self) -> None:
        if len(self._results) >= self._expected and self._results:
            out, self._results = self._results, []
            try:
                self._ex.transport.send(self._orig, {"type": "TableAccessResBatch", "results": out})
            except NetworkError:
                LOG.warning("origin %s of a request batch is gone", self._orig)

    def add(self, res: Dict[str, Any]) -> None:
        with self._lock:
            self._results.append(res)
            self._maybe_send()

    def drop(self) -> None:
        with self._lock:
            self._expected -= 1
            self._maybe_send()


class CommManager:
    """`num_threads` worker threads with bounded queues; an op is pinned to thread
    `block_id % num_threads`, which yields per-block FIFO execution."""

    def __init__(self, num_threads: int, queue_size: int, name: str) -> None:
        # `queue_size` is a soft limit (a warning): ops are enqueued by the executor's inbox thread, and an inbox that
        # blocks on a full queue can no longer deliver the ownership update the queued ops are waiting for
        self._queues = [queue.Queue() for _ in range(max(1, num_threads))]
        self._soft_limit, self._warned, self._name = max(1, queue_size), False, name
        self._threads = []
        self._closed = False
        for i, q in enumerate(self._queues):
            t = threading.Thread(target=self._run, args=(q,), name=f"{name}-comm-{i}", daemon=True)
            t.start()
            self._threads.append(t)

    def enqueue(self, block_id: int, fn: Callable[[], None]) -> None:
        q = self._queues[block_id % len(self._queues)]
        if not self._warned and q.qsize() >= self._soft_limit:
            self._warned = True
            LOG.warning("%s: more than %d table ops queued on one comm thread (slow or migrating block?)",
                        self._name, self._soft_limit)
        q.put(fn)

    def _run(self, q: "queue.Queue") -> None:
        while True:
            fn = q.get()
            if fn is None:
                return
            try:
                fn()
            except Exception:  # noqa: BLE001
                LOG.exception("comm op failed")

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        for q in self._queues:
            q.put(None)


def execute_on_block(block, op_type: str, keys, values, single: bool):
    """Run one op against a local Block; returns the per-key result list (or single value)."""
    if single:
        k = keys[0]
        v = None if values is None else values[0]
        if op_type == PUT:
            return [block.put(k, v)]
        if op_type == PUT_IF_ABSENT:
            return [block.put_if_absent(k, v)]
        if op_type == GET:
            return [block.get(k)]
        if op_type == GET_OR_INIT:
            return [block.get_or_init(k)]
        if op_type == REMOVE:
            return [block.remove(k)]
        if op_type == UPDATE:
            return [block.update(k, v)]
    else:
        if op_type == PUT:
            block.multi_put(keys, values)
            return None
        if op_type == GET:
            return list(block.multi_get(keys))
        if op_type == GET_OR_INIT:
            return list(block.multi_get_or_init(keys))
        if op_type == UPDATE:
            return list(block.multi_update(keys, values))
    raise ValueError(f"unsupported op {op_type} (single={single})")


class RemoteAccessOpSender:
    """Origin side: bookkeeping of in-flight ops + (re)sending to the owner."""

    def __init__(self, executor) -> None:
        self._ex = executor
        self._ongoing: Dict[int, RemoteDataOp] = {}
        self._lock = threading.Lock()
        self._flushed = threading.Condition(self._lock)
        self._op_ids = itertools.count()
        self.stat = RemoteAccessOpStat()

    def next_op_id(self) -> int:
        return next(self._op_ids)

    def num_ongoing(self) -> int:
        with self._lock:
            return len(self._ongoing)

    def send_op(self, op: RemoteDataOp, target_id: str) -> None:
        if op.reply_required:
            with self._lock:
                self._ongoing[op.op_id] = op
        if op.op_type in (GET, GET_OR_INIT):
            self.stat.inc_count_sent_get_req()
        self._send_with_retry(op, target_id)

    def register_redirected_local(self, op: RemoteDataOp) -> None:
        """A LOCAL op that bounced to the driver fallback (its table is being dropped / flushed here) comes back as a
        TableAccessRes from the real owner: it must be findable in `_ongoing`, or its caller waits for ever."""
        if op.reply_required:
            with self._lock:
                self._ongoing[op.op_id] = op

    def send_ops(self, ops: List[RemoteDataOp], target_id: str) -> None:
        """All block ops of one multi-key request that live on `target_id`, in one message (the reference sends one
        message per block: TableImpl.java:460-501; a GPU-era host has far fewer, fatter peers)."""
        if len(ops) == 1:
            return self.send_op(ops[0], target_id)
        with self._lock:
            for op in ops:
                if op.reply_required:
                    self._ongoing[op.op_id] = op
        if ops[0].op_type in (GET, GET_OR_INIT):
            self.stat.inc_count_sent_get_req()
        try:
            self._ex.transport.send(target_id, {"type": "TableAccessReqBatch", "orig_id": self._ex.executor_id,
                                                "ops": [op.to_msg() for op in ops]})
        except NetworkError:
            for op in ops:                      # stale ownership: fall back to the per-op path (re-resolves owners)
                self._send_with_retry(op, target_id)

    def on_table_access_res_batch(self, m: Dict[str, Any]) -> None:
        for res in m["results"]:
            self.on_table_access_res(res)

    def _send_with_retry(self, op: RemoteDataOp, target_id: str, max_retries: int = 100) -> None:
        comps = self._ex.tables.try_get_components(op.table_id)
        for _ in range(max_retries):
            try:
                self._ex.transport.send(target_id, op.to_msg())
                return
            except NetworkError:
                # stale ownership: ask the driver for the truth, back off, re-resolve, resend
                LOG.info("[%s] send of op %s to %s failed; re-resolving owner of block %s",
                         self._ex.executor_id, op.op_id, target_id, op.block_id)
                try:
                    self._ex.transport.send(self._ex.driver_id, {"type": "OwnershipReq", "table_id": op.table_id,
                                                                  "block_id": op.block_id, "src": self._ex.executor_id})
                except NetworkError:
                    pass
                time.sleep(RESEND_INTERVAL_S)
                if comps is None:
                    break
                # lock-free read of the ownership array: the caller may still hold this block's read lock (first
                # attempt of a multi-key request), and re-acquiring it behind a waiting ownership update would deadlock
                owner = comps.ownership.resolve_executor(op.block_id)
                if owner is None:   # became local meanwhile
                    self._ex.handler.enqueue_local(op)
                    return
                target_id = owner
        self._fail(op)

    def _fail(self, op: RemoteDataOp) -> None:
        with self._lock:
            self._ongoing.pop(op.op_id, None)
            self._flushed.notify_all()
        if op.callback is not None:
            op.callback(None, False)

    def on_table_access_res(self, m: Dict[str, Any]) -> None:
        with self._lock:
            op = self._ongoing.pop(m["op_id"], None)
            self._flushed.notify_all()
        if op is None:
            LOG.warning("response for unknown op %s", m["op_id"])
            return
        vals = m.get("values")
        if op.op_type in (GET, GET_OR_INIT) and vals is not None:
            nbytes = sum(getattr(v, "nbytes", 8) for v in vals)
            self.stat.inc_bytes_received_get_resp(nbytes)
        if op.callback is not None:
            op.callback(vals, m.get("success", True))

    def complete_local(self, op: RemoteDataOp, values, success: bool) -> None:
        if op.callback is not None:
            op.callback(values, success)

    def wait_ops_to_be_flushed(self, table_id: str, timeout: float = 60.0) -> None:
        deadline = time.time() + timeout
        with self._lock:
            while any(o.table_id == table_id for o in self._ongoing.values()):
                rem = deadline - time.time()
                if rem <= 0:
                    raise TimeoutError(f"ops of table {table_id} were not flushed")
                self._flushed.wait(rem)


class _OpTracker:
    """Per-table op tracker FSM: PROCESSING -> FLUSHING -> FLUSHED."""

    def __init__(self) -> None:
        self.sm = (StateMachine.new_builder()
                   .add_state("PROCESSING").add_state("FLUSHING").add_state("FLUSHED")
                   .add_transition("PROCESSING", "FLUSHING").add_transition("FLUSHING", "FLUSHED")
                   .add_transition("PROCESSING", "FLUSHED")
                   .set_initial_state("PROCESSING").build())
        self.in_flight = 0
        self.cond = threading.Condition()


class RemoteAccessOpHandler:
    """Owner side: executes ops on local blocks in per-block FIFO order."""

    def __init__(self, executor, num_threads: int, queue_size: int) -> None:
        self._ex = executor
        self._comm = CommManager(num_threads, queue_size, f"{executor.executor_id}-handler")
        self._trackers: Dict[str, _OpTracker] = {}
        self._lock = threading.Lock()

    def _tracker(self, table_id: str) -> _OpTracker:
        with self._lock:
            return self._trackers.setdefault(table_id, _OpTracker())

    def on_table_access_req(self, m: Dict[str, Any]) -> None:
        op = RemoteDataOp.from_msg(m)
        self._enqueue(op)

    def on_table_access_req_batch(self, m: Dict[str, Any]) -> None:
        subs = m["ops"]
        batch = _ReplyBatch(sum(1 for s in subs if s["reply_required"]), m["orig_id"], self._ex)
        for sub in subs:
            op = RemoteDataOp.from_msg(sub)
            if op.reply_required:
                op.batch = batch
            self._enqueue(op)

    def enqueue_local(self, op: RemoteDataOp) -> None:
        op.is_local = True
        self._enqueue(op)

    def execute_local(self, op: RemoteDataOp, comps) -> None:
        """Run an op on a LOCAL block in the caller's thread.  The caller holds the block's ownership read lock
        (TableImpl resolved the owner with it), so the block cannot migrate away underneath; blocks serialise
        their own mutations.  This is the reference's local path (TableImpl.java:366-408 calls the tablet directly);
        only requests that arrive over the network go through the per-block comm threads."""
        op.is_local = True
        tr = self._tracker(op.table_id)
        with tr.cond:
            if tr.sm.current_state != "PROCESSING":
                self._redirect_to_driver(op)
                return
            tr.in_flight += 1
        try:
            try:
                block = comps.block_store.get(op.block_id)
                result = execute_on_block(block, op.op_type, op.keys, op.values, op.single)
                ok = True
            except Exception:  # noqa: BLE001
                LOG.exception("local op %s on table %s failed", op.op_type, op.table_id)
                result, ok = None, False
            self._ex.sender.complete_local(op, result, ok)
        finally:
            with tr.cond:
                tr.in_flight -= 1
                tr.cond.notify_all()

    def _enqueue(self, op: RemoteDataOp) -> None:
        tr = self._tracker(op.table_id)
        with tr.cond:
            if tr.sm.current_state != "PROCESSING":
                # table is being dropped here: let the driver re-route (fallback)
                self._redirect_to_driver(op)
                return
            tr.in_flight += 1
        self._comm.enqueue(op.block_id, lambda: self._process(op, tr))

    def _process(self, op: RemoteDataOp, tr: _OpTracker) -> None:
        try:
            comps = self._ex.tables.try_get_components(op.table_id)
            if comps is None:
                self._redirect_to_driver(op)
                return
            owner, unlock = comps.ownership.resolve_executor_with_lock(op.block_id)
            try:
                if owner is not None:
                    # stale request: forward to the real owner, keeping the original origin
                    self._redirect(op, owner)
                    return
                block = comps.block_store.get(op.block_id)
                try:
                    result = execute_on_block(block, op.op_type, op.keys, op.values, op.single)
                    ok = True
                except Exception:  # noqa: BLE001
                    LOG.exception("op %s on table %s failed", op.op_type, op.table_id)
                    result, ok = None, False
            finally:
                unlock()
            self._reply(op, result, ok)
        finally:
            with tr.cond:
                tr.in_flight -= 1
                tr.cond.notify_all()

    def _reply(self, op: RemoteDataOp, result, ok: bool) -> None:
        if op.is_local or op.orig_id == self._ex.executor_id:
            if op.is_local:
                self._ex.sender.complete_local(op, result, ok)
            elif op.reply_required:
                self._ex.sender.on_table_access_res({"op_id": op.op_id, "values": result, "success": ok})
            return
        if not op.reply_required:
            return
        res = {"type": "TableAccessRes", "op_id": op.op_id, "table_id": op.table_id, "block_id": op.block_id,
               "values": result, "success": ok, "single": op.single}
        if op.batch is not None:
            op.batch.add(res)
            return
        try:
            self._ex.transport.send(op.orig_id, res)
        except NetworkError:
            LOG.warning("origin %s of op %s is gone", op.orig_id, op.op_id)

    def _redirect(self, op: RemoteDataOp, owner: str) -> None:
        if op.is_local:
            # a local op whose block moved away between resolve and execution
            op.is_local = False
            op.orig_id = self._ex.executor_id
            self._ex.sender.send_op(op, owner)
            return
        self._forward(op, owner)

    def _leave_batch(self, op: RemoteDataOp) -> None:
        if op.batch is not None:
            op.batch.drop()
            op.batch = None

    def _forward(self, op: RemoteDataOp, owner: str) -> None:
        self._leave_batch(op)
        try:
            self._ex.transport.send(owner, op.to_msg())
        except NetworkError:
            self._redirect_to_driver(op)

    def _redirect_to_driver(self, op: RemoteDataOp) -> None:
        self._leave_batch(op)
        msg = op.to_msg()
        msg["fallback"] = True
        if op.is_local:
            msg["orig_id"] = self._ex.executor_id
            self._ex.sender.register_redirected_local(op)
        try:
            self._ex.transport.send(self._ex.driver_id, msg)
        except NetworkError:
            LOG.error("cannot reach the driver for fallback of op %s", op.op_id)
            if op.is_local:
                self._ex.sender._fail(op)       # fail the caller now instead of hanging it

    def wait_ops_to_be_flushed(self, table_id: str, timeout: float = 60.0) -> None:
        tr = self._tracker(table_id)
        with tr.cond:
            if tr.sm.current_state == "PROCESSING":
                tr.sm.set_state("FLUSHING")
            if not tr.cond.wait_for(lambda: tr.in_flight == 0, timeout):
                raise TimeoutError(f"handler ops of table {table_id} were not flushed")
            if tr.sm.current_state == "FLUSHING":
                tr.sm.set_state("FLUSHED")

    def forget_table(self, table_id: str) -> None:
        with self._lock:
            self._trackers.pop(table_id, None)

    def close(self) -> None:
        self._comm.close()
