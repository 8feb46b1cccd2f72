"""Driver (master) side of the Elastic Table service.

One `ETMaster` lives in the job-server process.  It mirrors the reference's driver components:

  ETMasterImpl / ExecutorManager / AllocatedExecutorImpl   driver/impl/ETMasterImpl.java:44-88,
                                                            ExecutorManager.java:122-243
  TableManager / AllocatedTableImpl                         TableManager.java:75-157,
                                                            AllocatedTableImpl.java:95-395
  BlockManager                                              BlockManager.java:76-88,156-245
  MigrationManager                                          MigrationManager.java:76-172
  SubscriptionManager                                       SubscriptionManager.java:67-207
  TableControlAgent                                         TableControlAgent.java:66-237
  ChkpManagerMaster                                         ChkpManagerMaster.java:142-428
  FallbackManager                                           FallbackManager.java:55-97
  GlobalTaskUnitScheduler                                   GlobalTaskUnitScheduler.java:52-92
  RunningTasklet / TaskletRepresenter                       TaskletRepresenter.java:41-129
  MetricManager / MetricReceiver                            metric/MetricManager.java:46-56

Executors are obtained through an `ExecutorLauncher` (the EvaluatorManager equivalent,
services/evaluator-manager/.../EvaluatorManager.java:39-86): in-process threads, or one OS
process per GPU.
"""
from __future__ import annotations

import itertools
import logging
import os
import threading
import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Set, Tuple

from ..utils.concurrent import AggregateFuture, CompletedFuture, ResultFuture, new_fixed_thread_pool
from ..utils.state_machine import StateMachine
from . import checkpoint as chk
from .config import ExecutorConfiguration, TableConfiguration, TaskletConfiguration
from .exceptions import (AlreadyAssociatedError, ChkpNotExistError, ExecutorNotExistError, NotAssociatedError, TableNotExistError)
from .transport import NetworkError, Transport

LOG = logging.getLogger(__name__)
NUM_SPLITS_PER_EXECUTOR = 32   # TableControlAgent.java:43


# ==========================================================================================
# Block manager
# ==========================================================================================
class BlockManager:
    """Driver truth of block -> owner, per-executor block sets and the set of moving blocks."""

    def __init__(self, num_total_blocks: int) -> None:
        self.num_total_blocks = num_total_blocks
        self._owner: List[Optional[str]] = [None] * num_total_blocks
        self._blocks_of: Dict[str, Set[int]] = {}
        self._moving: Set[int] = set()
        self._lock = threading.RLock()

    def init(self, executor_ids: Sequence[str]) -> None:
        """Round-robin initial placement: block b -> executors[b % n]."""
        ids = sorted(executor_ids)
        with self._lock:
            for e in ids:
                self._blocks_of[e] = set()
            for b in range(self.num_total_blocks):
                e = ids[b % len(ids)]
                self._owner[b] = e
                self._blocks_of[e].add(b)

    def register_executor(self, executor_id: str) -> None:
        with self._lock:
            if executor_id in self._blocks_of:
                raise AlreadyAssociatedError(executor_id)
            self._blocks_of[executor_id] = set()

    def deregister_executor(self, executor_id: str) -> None:
        with self._lock:
            if self._blocks_of.get(executor_id):
                raise RuntimeError(f"executor {executor_id} still owns blocks")
            self._blocks_of.pop(executor_id, None)

    def get_associator_ids(self) -> Set[str]:
        with self._lock:
            return set(self._blocks_of)

    def get_partition_info(self) -> Dict[str, Set[int]]:
        with self._lock:
            return {e: set(bs) for e, bs in self._blocks_of.items()}

    def get_ownership_status(self) -> List[str]:
        with self._lock:
            return list(self._owner)

    def get_owner_id(self, block_id: int) -> str:
        with self._lock:
            return self._owner[block_id]

    def get_num_blocks(self, executor_id: str) -> int:
        with self._lock:
            return len(self._blocks_of.get(executor_id, ()))

    def choose_blocks_to_move(self, src: str, num_blocks: int) -> List[int]:
        with self._lock:
            if src not in self._blocks_of:
                raise NotAssociatedError(src)
            cands = sorted(self._blocks_of[src] - self._moving)[:num_blocks]
            self._moving.update(cands)
            return cands

    def update_owner(self, block_id: int, old: str, new: str) -> None:
        with self._lock:
            if self._owner[block_id] != old:
                raise RuntimeError(f"block {block_id}: owner is {self._owner[block_id]}, expected {old}")
            if block_id not in self._moving:
                raise RuntimeError(f"block {block_id} was not chosen for migration")
            self._owner[block_id] = new
            self._blocks_of[old].discard(block_id)
            self._blocks_of.setdefault(new, set()).add(block_id)

    def release_block_from_move(self, block_id: int) -> None:
        with self._lock:
            if block_id not in self._moving:
                raise RuntimeError(f"block {block_id} is not moving")
            self._moving.discard(block_id)

    def num_moving(self) -> int:
        with self._lock:
            return len(self._moving)


# ==========================================================================================
# Subscription manager
# ==========================================================================================
class SubscriptionManager:
    """Which executors hold an ownership cache of a table; broadcasts ownership updates and
    replays the ones that happened while a new subscriber was initialising."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self._subs: Dict[str, Set[str]] = {}
        self._pending: Dict[int, Dict[str, Any]] = {}
        self._op_ids = itertools.count()
        self._lock = threading.RLock()

    def register_subscription(self, table_id: str, executor_id: str) -> None:
        with self._lock:
            self._subs.setdefault(table_id, set()).add(executor_id)

    def unregister_subscription(self, table_id: str, executor_id: str) -> None:
        with self._lock:
            self._subs.get(table_id, set()).discard(executor_id)

    def unregister_subscribers(self, table_id: str) -> Set[str]:
        with self._lock:
            return self._subs.pop(table_id, set())

    def get_subscribers(self, table_id: str) -> Set[str]:
        with self._lock:
            return set(self._subs.get(table_id, ()))

    def start_subscription_init(self, table_id: str, executor_ids: Set[str]) -> int:
        with self._lock:
            op = next(self._op_ids)
            self._pending[op] = {"table_id": table_id, "executors": set(executor_ids), "updates": []}
            return op

    def finish_subscription_init(self, op_id: int) -> None:
        with self._lock:
            p = self._pending.pop(op_id)
            for e in p["executors"]:
                self._subs.setdefault(p["table_id"], set()).add(e)
            updates = list(p["updates"])
        for (block_id, old, new) in updates:   # replay
            for e in p["executors"]:
                self._m.send(e, {"type": "OwnershipUpdate", "table_id": p["table_id"], "block_id": block_id,
                                 "old": self._m.driver_id, "new": new})

    def broadcast_update(self, table_id: str, block_id: int, old: str, new: str) -> int:
        """To every subscriber except the migration's sender and receiver. Returns #messages."""
        with self._lock:
            targets = set(self._subs.get(table_id, ())) - {old, new}
            for p in self._pending.values():
                if p["table_id"] == table_id:
                    p["updates"].append((block_id, old, new))
        for e in targets:
            self._m.send(e, {"type": "OwnershipUpdate", "table_id": table_id, "block_id": block_id,
                             "old": old, "new": new})
        return len(targets)


# ==========================================================================================
# Migration manager
# ==========================================================================================
class MigrationResult:
    """Outcome of a moveBlocks call.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/MigrationResult.java."""

    def __init__(self, completed: bool, msg: str, migrated_blocks: List[int]) -> None:
        self.is_completed, self.msg, self.migrated_blocks = completed, msg, migrated_blocks


class _Migration:
    def __init__(self, op_id, table_id, src, dst, blocks, block_manager, future) -> None:
        self.op_id, self.table_id, self.src, self.dst = op_id, table_id, src, dst
        self.blocks, self.bm, self.future = list(blocks), block_manager, future
        self.owner_moved: Set[int] = set()
        self.data_moved: Set[int] = set()
        self.early_data_moved: Set[int] = set()
        self.lock = threading.Lock()


class MigrationManager:
    """Driver side of host-table migration: op ids, MoveInit to the sender, ownership / data moved bookkeeping
    (arrival order fixed per block), broadcast to the other subscribers, completion.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/MigrationManager.java:76-92,101-172."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self._ongoing: Dict[int, _Migration] = {}
        self._op_ids = itertools.count()
        self._lock = threading.Lock()

    def start_migration(self, block_manager: BlockManager, table_id: str, src: str, dst: str,
                        blocks: List[int]) -> ResultFuture:
        fut: ResultFuture = ResultFuture()
        with self._lock:
            op = next(self._op_ids)
            self._ongoing[op] = _Migration(op, table_id, src, dst, blocks, block_manager, fut)
        self._m.send(src, {"type": "MoveInit", "op_id": op, "table_id": table_id, "block_ids": list(blocks),
                           "sender_id": src, "receiver_id": dst})
        return fut

    def on_ownership_moved(self, op_id: int, block_id: int) -> None:
        with self._lock:
            mig = self._ongoing.get(op_id)
        if mig is None:
            return
        with mig.lock:
            mig.bm.update_owner(block_id, mig.src, mig.dst)
            mig.owner_moved.add(block_id)
            replay = block_id in mig.early_data_moved
        self._m.subscriptions.broadcast_update(mig.table_id, block_id, mig.src, mig.dst)
        if replay:   # DataMoved overtook OwnershipMoved (reordering tolerated like the reference)
            self._finish_block(mig, block_id)

    def on_data_moved(self, op_id: int, block_id: int, ownership_together: bool) -> None:
        with self._lock:
            mig = self._ongoing.get(op_id)
        if mig is None:
            return
        if ownership_together:
            with mig.lock:
                mig.bm.update_owner(block_id, mig.src, mig.dst)
                mig.owner_moved.add(block_id)
            self._m.subscriptions.broadcast_update(mig.table_id, block_id, mig.src, mig.dst)
        else:
            with mig.lock:
                if block_id not in mig.owner_moved:
                    mig.early_data_moved.add(block_id)
                    return
        self._finish_block(mig, block_id)

    def _finish_block(self, mig: _Migration, block_id: int) -> None:
        with mig.lock:
            if block_id in mig.data_moved:
                return
            mig.data_moved.add(block_id)
            mig.bm.release_block_from_move(block_id)
            done = len(mig.data_moved) == len(mig.blocks)
        if done:
            with self._lock:
                self._ongoing.pop(mig.op_id, None)
            mig.future.on_completed(MigrationResult(True, "migration completed", list(mig.blocks)))


# ==========================================================================================
# Table control agent (fan-out / fan-in)
# ==========================================================================================
class TableControlAgent:
    """Fan-out / fan-in of table control messages (init, load with 32 splits per executor, drop, ownership sync,
    device peer map / quiesce / flip) with aggregate futures; a failed executor fails the future.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/TableControlAgent.java:43,66-133,163-237."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self._futs: Dict[int, AggregateFuture] = {}
        self._op_ids = itertools.count()
        self._lock = threading.Lock()

    def _new(self, n: int) -> Tuple[int, AggregateFuture]:
        fut = AggregateFuture(n)
        with self._lock:
            op = next(self._op_ids)
            self._futs[op] = fut
        return op, fut

    def on_ack(self, op_id: int, payload: Any = None) -> None:
        with self._lock:
            fut = self._futs.get(op_id)
        if fut is not None:
            fut.on_partial(payload)
            if fut.done():
                with self._lock:
                    self._futs.pop(op_id, None)

    def on_failed(self, op_id: int, error: str) -> None:
        """An executor could not perform its part of control op `op_id`: fail the whole fan-out."""
        with self._lock:
            fut = self._futs.pop(op_id, None)
        if fut is not None:
            fut.on_failed(RuntimeError(error))

    def init_table(self, conf: TableConfiguration, executor_ids: Set[str], owners: List[str]) -> AggregateFuture:
        op, fut = self._new(len(executor_ids))
        for e in executor_ids:
            self._m.send(e, {"type": "TableInit", "op_id": op, "table_conf": conf.to_dict(), "block_owners": owners})
        return fut

    def load(self, table_id: str, executor_ids: Set[str], input_path: str) -> AggregateFuture:
        from ..common.dataloader import get_splits
        ids = sorted(executor_ids)
        try:
            splits = get_splits(input_path, NUM_SPLITS_PER_EXECUTOR * len(ids))
        except OSError as e:
            failed: AggregateFuture = AggregateFuture(len(ids))
            failed.on_failed(RuntimeError(f"cannot read input '{input_path}' of table {table_id}: {e}"))
            return failed
        op, fut = self._new(len(ids))
        for i, e in enumerate(ids):
            mine = splits[i * NUM_SPLITS_PER_EXECUTOR:(i + 1) * NUM_SPLITS_PER_EXECUTOR]
            self._m.send(e, {"type": "TableLoad", "op_id": op, "table_id": table_id,
                             "file_splits": [s.to_dict() for s in mine]})
        return fut

    def drop_table(self, table_id: str, executor_ids: Set[str]) -> AggregateFuture:
        op, fut = self._new(len(executor_ids))
        for e in executor_ids:
            try:
                self._m.send(e, {"type": "TableDrop", "op_id": op, "table_id": table_id})
            except NetworkError:
                fut.on_partial(None)
        return fut

    def fan_out(self, executor_ids: Set[str], msg: Dict[str, Any]) -> AggregateFuture:
        """Send `msg` (+ op_id) to every executor; the future completes with all the acks."""
        op, fut = self._new(len(executor_ids))
        for e in executor_ids:
            self._m.send(e, dict(msg, op_id=op))
        return fut

    def sync_ownership(self, table_id: str, deleted_executor_id: str, executor_ids: Set[str]) -> AggregateFuture:
        op, fut = self._new(len(executor_ids))
        for e in executor_ids:
            self._m.send(e, {"type": "OwnershipSync", "op_id": op, "table_id": table_id,
                             "deleted_executor_id": deleted_executor_id})
        return fut


# ==========================================================================================
# Checkpoint master
# ==========================================================================================
class ChkpManagerMaster:
    """Checkpoint ids `<tableId>-<ms>`, ChkpStart to the associators, temp vs committed blocks, restore (possibly
    with a different blocking / executor set) from id or path.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/ChkpManagerMaster.java:142-168,177-206,208-245,250-301,309-418."""

    def __init__(self, master: "ETMaster", temp_path: str, commit_path: str) -> None:
        self._m = master
        self.temp_path, self.commit_path = temp_path, commit_path
        self._chkps: Dict[str, Dict[str, Any]] = {}
        self._load_futs: Dict[int, AggregateFuture] = {}
        self._op_ids = itertools.count()
        self._lock = threading.Lock()

    def checkpoint(self, conf: TableConfiguration, associators: Set[str], sampling_ratio: float) -> ResultFuture:
        chkp_id = f"{conf.table_id}-{int(time.time() * 1000)}"
        fut: ResultFuture = ResultFuture()
        with self._lock:
            while chkp_id in self._chkps:
                chkp_id += "x"
            self._chkps[chkp_id] = {"conf": conf, "remaining": set(associators), "temp_blocks": {},
                                    "committed": set(), "future": fut, "sampling": sampling_ratio}
        for e in associators:
            self._m.send(e, {"type": "ChkpStart", "chkp_id": chkp_id, "table_id": conf.table_id,
                             "sampling_ratio": sampling_ratio})
        return fut

    def on_chkp_done(self, chkp_id: str, executor_id: str, block_ids: List[int]) -> None:
        with self._lock:
            c = self._chkps.get(chkp_id)
            if c is None:             # the checkpoint already failed on another executor
                return
            c["temp_blocks"][executor_id] = list(block_ids)
            c["remaining"].discard(executor_id)
            done = not c["remaining"]
        if done:
            c["future"].on_completed(chkp_id)

    def on_chkp_failed(self, m: Dict[str, Any]) -> None:
        """An executor could not write (ChkpStart) or restore (ChkpLoad, has an op_id) its part: fail the waiting
        future instead of letting `checkpoint().get()` / `create_table_from_chkp().get()` run into their timeouts."""
        err = RuntimeError(f"checkpoint {m['chkp_id']} failed on {m.get('executor_id')}: {m.get('error')}")
        LOG.error("%s", err)
        if "op_id" in m:
            with self._lock:
                fut = self._load_futs.pop(m["op_id"], None)
            if fut is not None and not fut.done():
                fut.on_failed(err)
            return
        with self._lock:
            c = self._chkps.pop(m["chkp_id"], None)
        if c is not None and not c["future"].done():
            c["future"].on_failed(err)

    def on_commit_ack(self, chkp_id: str, executor_id: str) -> None:
        with self._lock:
            c = self._chkps.get(chkp_id)
            if c is not None:
                c["committed"].update(c["temp_blocks"].pop(executor_id, []))

    def on_executor_closed(self, executor_id: str) -> None:
        """Executors commit their temp checkpoints when they stop (ContextStopHandler)."""
        with self._lock:
            for c in self._chkps.values():
                c["committed"].update(c["temp_blocks"].pop(executor_id, []))

    def commit(self, chkp_id: str) -> None:
        with self._lock:
            c = self._chkps.get(chkp_id)
            targets = list(c["temp_blocks"]) if c else []
        for e in targets:
            self._m.send(e, {"type": "ChkpCommit", "chkp_id": chkp_id})

    def get_table_conf(self, chkp_id: str) -> TableConfiguration:
        with self._lock:
            c = self._chkps.get(chkp_id)
        if c is not None:
            return c["conf"]
        d = os.path.join(self.commit_path, chkp_id)
        if os.path.isdir(d):
            return chk.read_conf(d)
        raise ChkpNotExistError(chkp_id)

    def load(self, chkp_id: str, table: "AllocatedTable") -> ResultFuture:
        """Restore into `table` (which may have different blocks / executors than the source)."""
        with self._lock:
            c = self._chkps.get(chkp_id)
        assoc = sorted(table.block_manager.get_associator_ids())
        plan: Dict[str, Dict[str, Any]] = {}
        if c is not None:
            # blocks still in an executor's temp dir are loaded by that executor when alive
            for e, blocks in c["temp_blocks"].items():
                tgt = e if e in assoc else assoc[hash(e) % len(assoc)]
                plan.setdefault(tgt, {"blocks": [], "committed": False})["blocks"].extend(blocks)
            committed = sorted(c["committed"])
        else:
            d = os.path.join(self.commit_path, chkp_id)
            if not os.path.isdir(d):
                raise ChkpNotExistError(chkp_id)
            committed = chk.list_chkp_blocks(d)
        loads: List[Tuple[str, List[int], bool]] = [(e, p["blocks"], False) for e, p in plan.items()]
        for i, e in enumerate(assoc):
            mine = committed[i::len(assoc)]
            if mine:
                loads.append((e, mine, True))
        return self._dispatch_loads(chkp_id, table.table_id, loads)

    def load_from_path(self, path: str, table: "AllocatedTable") -> ResultFuture:
        from ..common.dataloader import strip_scheme
        path = strip_scheme(path)
        blocks = chk.list_chkp_blocks(path)
        assoc = sorted(table.block_manager.get_associator_ids())
        loads = [(e, blocks[i::len(assoc)], True) for i, e in enumerate(assoc) if blocks[i::len(assoc)]]
        return self._dispatch_loads(path, table.table_id, loads)

    def _dispatch_loads(self, chkp_id: str, table_id: str, loads) -> ResultFuture:
        fut = AggregateFuture(len(loads))
        with self._lock:
            op = next(self._op_ids)
            self._load_futs[op] = fut
        for e, blocks, committed in loads:
            self._m.send(e, {"type": "ChkpLoad", "op_id": op, "chkp_id": chkp_id, "table_id": table_id,
                             "block_ids": list(blocks), "committed": committed})
        return fut

    def on_load_done(self, op_id: int, payload: Any) -> None:
        with self._lock:
            fut = self._load_futs.get(op_id)
        if fut is not None:
            fut.on_partial(payload)


# ==========================================================================================
# Fallback manager
# ==========================================================================================
class FallbackManager:
    """Re-routes table access requests that bounced (executor gone / table dropped there) to
    the current owner, retrying every 100 ms while the owner is unreachable."""
    RETRY_S = 0.1

    def __init__(self, master: "ETMaster") -> None:
        self._m = master

    def on_table_access_req(self, m: Dict[str, Any]) -> None:
        m = dict(m)
        m.pop("fallback", None)
        for _ in range(100):
            try:
                table = self._m.table_manager.get_allocated_table(m["table_id"])
            except TableNotExistError:
                LOG.warning("fallback: table %s does not exist; dropping op %s", m["table_id"], m["op_id"])
                self._fail(m)
                return
            owner = table.block_manager.get_owner_id(m["block_id"])
            try:
                self._m.send(owner, m)
                return
            except NetworkError:
                time.sleep(self.RETRY_S)
        self._fail(m)

    def _fail(self, m: Dict[str, Any]) -> None:
        if m.get("reply_required"):
            try:
                self._m.send(m["orig_id"], {"type": "TableAccessRes", "op_id": m["op_id"], "table_id": m["table_id"],
                                            "block_id": m["block_id"], "values": None, "success": False,
                                            "single": m["single"]})
            except NetworkError:
                pass


# ==========================================================================================
# Global task unit scheduler
# ==========================================================================================
class GlobalTaskUnitScheduler:
    """Per (job, task unit, seq) barrier: Ready is broadcast when every executor of the job waits
    on the same unit, so the phases of different jobs interleave on shared executors."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self._job_size: Dict[str, int] = {}
        self._waiting: Dict[Tuple[str, str, int], List[Tuple[str, str]]] = {}
        self._lock = threading.Lock()
        self.num_waits = 0                 # TaskUnitWait messages served (one driver round trip each)

    def on_job_start(self, job_id: str, num_executors: int) -> None:
        with self._lock:
            self._job_size[job_id] = num_executors

    def on_job_finish(self, job_id: str) -> None:
        with self._lock:
            self._job_size.pop(job_id, None)
            for k in [k for k in self._waiting if k[0] == job_id]:      # barriers a failed job never completed
                self._waiting.pop(k, None)

    def on_job_resize(self, job_id: str, num_executors: int) -> None:
        with self._lock:
            self._job_size[job_id] = num_executors
            ready = [k for k, v in self._waiting.items() if k[0] == job_id and len(v) >= num_executors]
            release = [(k, self._waiting.pop(k)) for k in ready]
        for k, waiters in release:
            self._release(k, waiters)

    def on_leave(self, job_id: str, executor_id: str) -> None:
        """A worker finished its loop: it no longer takes part in the job's unit barriers."""
        with self._lock:
            n = self._job_size.get(job_id)
        if n is not None and n > 1:
            self.on_job_resize(job_id, n - 1)

    def on_join(self, job_id: str, executor_id: str) -> None:
        with self._lock:
            if job_id in self._job_size:
                self._job_size[job_id] += 1

    def on_wait(self, m: Dict[str, Any]) -> None:
        key = (m["job_id"], m["unit_name"], m["seq"])
        with self._lock:
            self.num_waits += 1
            n = self._job_size.get(m["job_id"], 1)
            lst = self._waiting.setdefault(key, [])
            lst.append((m["executor_id"], m["tasklet_id"]))
            if len(lst) < n:
                return
            waiters = self._waiting.pop(key)
        self._release(key, waiters)

    def _release(self, key, waiters) -> None:
        for ex, tasklet_id in waiters:
            self._m.send(ex, {"type": "TaskUnitReady", "tasklet_id": tasklet_id, "job_id": key[0],
                              "unit_name": key[1], "seq": key[2], "inline": True})


# ==========================================================================================
# Tasklets (driver view)
# ==========================================================================================
class TaskletResult:
    """Final status + result of a tasklet.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/TaskletResult.java."""

    def __init__(self, tasklet_id: str, executor_id: str, status: str, result: Any, error: Optional[str]) -> None:
        self.tasklet_id, self.executor_id, self.status, self.result, self.error = \
            tasklet_id, executor_id, status, result, error

    @property
    def is_success(self) -> bool:
        return self.status == "Done"


class RunningTasklet:
    """Driver handle (FSM INIT -> RUNNING -> DONE | FAILED; Done before Running tolerated)."""

    def __init__(self, master: "ETMaster", executor_id: str, conf: TaskletConfiguration,
                 custom_handler: Optional[Callable[[Any], None]] = None) -> None:
        self._m = master
        self.executor_id = executor_id
        self.tasklet_id = conf.tasklet_id
        self.custom_handler = custom_handler
        self.sm = (StateMachine.new_builder().add_state("INIT").add_state("RUNNING").add_state("DONE")
                   .add_state("FAILED").add_transition("INIT", "RUNNING").add_transition("RUNNING", "DONE")
                   .add_transition("RUNNING", "FAILED").add_transition("INIT", "DONE").add_transition("INIT", "FAILED")
                   .set_initial_state("INIT").build())
        self._result: ResultFuture = ResultFuture()

    def on_status(self, m: Dict[str, Any]) -> None:
        st = m["status"]
        if st == "Running":
            self.sm.compare_and_set_state("INIT", "RUNNING")
        elif st in ("Done", "Failed"):
            self.sm.set_state("DONE" if st == "Done" else "FAILED")
            self._result.on_completed(TaskletResult(self.tasklet_id, self.executor_id, st, m.get("result"),
                                                    m.get("error")))

    def stop(self) -> None:
        self._m.send(self.executor_id, {"type": "TaskletStop", "tasklet_id": self.tasklet_id})

    def send(self, payload: Any, inline: bool = False) -> None:
        """`inline`: the tasklet's handler for this payload never blocks (a barrier release): the executor runs it on
        its inbox thread instead of handing it to the handler pool -- one thread hop less per mini-batch."""
        m = {"type": "TaskletCustom", "tasklet_id": self.tasklet_id, "payload": payload}
        if inline:
            m["inline"] = True
        self._m.send(self.executor_id, m)

    def get_task_result(self, timeout: Optional[float] = None) -> TaskletResult:
        return self._result.get(timeout)

    def is_done(self) -> bool:
        return self._result.done()


# ==========================================================================================
# Executors
# ==========================================================================================
class ExecutorLauncher:
    """SPI that produces executors (the EvaluatorManager equivalent)."""

    def launch(self, executor_id: str, conf: ExecutorConfiguration, master: "ETMaster") -> None:
        raise NotImplementedError

    def terminate(self, executor_id: str) -> None:
        pass


class InProcessLauncher(ExecutorLauncher):
    """Executors are objects in this process sharing the (Local)Transport."""

    def __init__(self, executor_factory: Optional[Callable[..., Any]] = None) -> None:
        self.executors: Dict[str, Any] = {}
        self._factory = executor_factory

    def launch(self, executor_id, conf, master):
        from .executor import Executor
        factory = self._factory or Executor
        self.executors[executor_id] = factory(executor_id, master.transport, conf, master.driver_id)
        master.on_executor_registered(executor_id)

    def terminate(self, executor_id):
        ex = self.executors.pop(executor_id, None)
        if ex is not None:
            ex.shutdown()


class AllocatedExecutor:
    """Driver handle of one executor: submit tasklets, close, running tasklets.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/AllocatedExecutorImpl.java:63-121."""

    def __init__(self, master: "ETMaster", executor_id: str, conf: ExecutorConfiguration) -> None:
        self._m = master
        self.id = executor_id
        self.conf = conf
        self._tasklets: Dict[str, RunningTasklet] = {}
        self._closed: ResultFuture = ResultFuture()
        self._lock = threading.Lock()

    def get_id(self) -> str:
        return self.id

    def submit_tasklet(self, conf: TaskletConfiguration,
                       custom_handler: Optional[Callable[[Any], None]] = None) -> RunningTasklet:
        rt = RunningTasklet(self._m, self.id, conf, custom_handler)
        with self._lock:
            self._tasklets[conf.tasklet_id] = rt
        self._m.register_tasklet(rt)
        self._m.send(self.id, {"type": "TaskletStart", "conf": conf.to_dict()})
        return rt

    def forget_tasklet(self, tasklet_id: str) -> None:
        with self._lock:
            self._tasklets.pop(tasklet_id, None)

    def get_running_tasklets(self) -> Dict[str, RunningTasklet]:
        with self._lock:
            return {k: v for k, v in self._tasklets.items() if not v.is_done()}

    def close(self) -> ResultFuture:
        try:
            self._m.send(self.id, {"type": "ExecutorClose"})
        except NetworkError:
            self._closed.on_completed(None)
        return self._closed

    def on_closed(self) -> None:
        self._closed.on_completed(None)


# ==========================================================================================
# Tables (driver view)
# ==========================================================================================
class AllocatedTable:
    """Lifecycle FSM UNINITIALIZED -> INITIALIZED -> DROPPED."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self.sm = (StateMachine.new_builder().add_state("UNINITIALIZED").add_state("INITIALIZED")
                   .add_state("DROPPED").add_transition("UNINITIALIZED", "INITIALIZED")
                   .add_transition("INITIALIZED", "DROPPED").set_initial_state("UNINITIALIZED").build())
        self.conf: Optional[TableConfiguration] = None
        self.block_manager: Optional[BlockManager] = None
        self._cond = threading.Condition()
        self._ongoing_migrations = 0
        self._ongoing_chkps = 0
        self.device_infos: Dict[str, Any] = {}     # executor id -> slab locations (device backend)

    @property
    def is_device(self) -> bool:
        return bool(self.conf is not None and self.conf.dense and self.conf.backend == "device")

    def _exchange_peer_map(self, acks: List[Dict[str, Any]], targets: Set[str]) -> ResultFuture:
        """Device backend, phase 2 of table init: ship every owner's slab map to `targets`."""
        for a in acks:
            if a and a.get("device_info") is not None:
                self.device_infos[a["src"]] = a["device_info"]
        return self._m.control.fan_out(targets, {"type": "TablePeerMap", "table_id": self.table_id,
                                                 "infos": list(self.device_infos.values())})

    @property
    def table_id(self) -> str:
        return self.conf.table_id

    def get_id(self) -> str:
        return self.conf.table_id

    # ---- init ----
    def init(self, conf: TableConfiguration, initial_associators: Sequence[AllocatedExecutor]) -> ResultFuture:
        self.sm.check_state("UNINITIALIZED")
        self.conf = conf
        self.block_manager = BlockManager(conf.num_total_blocks)
        ids = {e.id for e in initial_associators}
        for e in ids:
            self._m.subscriptions.register_subscription(conf.table_id, e)
        self.block_manager.init(sorted(ids))
        fut: ResultFuture = ResultFuture()
        init_fut = self._m.control.init_table(conf, ids, self.block_manager.get_ownership_status())

        def done(_=None):
            self.sm.set_state("INITIALIZED")
            fut.on_completed(self)

        def after_map(_=None):
            if conf.chkp_path:
                self._m.chkp_master.load_from_path(conf.chkp_path, self).chain(fut, done)
            elif conf.input_path:
                self._m.control.load(conf.table_id, self._loaders(ids), conf.input_path).chain(fut, done)
            else:
                done()

        def after_init(acks):
            if self.is_device:
                self._exchange_peer_map(acks, ids).chain(fut, after_map)
            else:
                after_map()
        init_fut.chain(fut, after_init)
        return fut

    def init_from_chkp(self, chkp_id: str, initial_associators: Sequence[AllocatedExecutor],
                       table_id: Optional[str] = None) -> ResultFuture:
        self.sm.check_state("UNINITIALIZED")
        conf = self._m.chkp_master.get_table_conf(chkp_id)
        if table_id is not None:
            conf = TableConfiguration.from_dict({**conf.to_dict(), "table_id": table_id})
        conf.chkp_path = None
        conf.input_path = None
        self.conf = conf
        self.block_manager = BlockManager(conf.num_total_blocks)
        ids = {e.id for e in initial_associators}
        for e in ids:
            self._m.subscriptions.register_subscription(conf.table_id, e)
        self.block_manager.init(sorted(ids))
        fut: ResultFuture = ResultFuture()

        def after_init(acks):
            def done(_=None):
                self.sm.set_state("INITIALIZED")
                fut.on_completed(self)

            def load(_=None):
                self._m.chkp_master.load(chkp_id, self).chain(fut, done)
            if self.is_device:
                self._exchange_peer_map(acks, ids).chain(fut, load)
            else:
                load()
        self._m.control.init_table(conf, ids, self.block_manager.get_ownership_status()).chain(fut, after_init)
        return fut

    def _loaders(self, ids: Set[str]) -> Set[str]:
        """The executors that read the input splits: those that own a block (a table with fewer blocks than
        associators leaves some without one, and the none-key loader draws its keys from the local blocks)."""
        owners = {e for e in ids if self.block_manager.get_num_blocks(e) > 0}
        return owners or set(ids)

    def load(self, executors: Sequence[AllocatedExecutor], input_path: str) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        return self._m.control.load(self.table_id, self._loaders({e.id for e in executors}), input_path)

    # ---- subscription / association ----
    def subscribe(self, executors: Sequence[AllocatedExecutor]) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        ids = {e.id for e in executors}
        op = self._m.subscriptions.start_subscription_init(self.table_id, ids)
        fut = self._m.control.init_table(self.conf, ids, self.block_manager.get_ownership_status())
        if self.is_device:
            out: ResultFuture = ResultFuture()

            def after(acks):
                def fin(_):
                    self._m.subscriptions.finish_subscription_init(op)
                    out.on_completed(None)
                self._exchange_peer_map(acks, ids).chain(out, fin)
            fut.chain(out, after)
            return out
        fut.add_listener(lambda _: self._m.subscriptions.finish_subscription_init(op))
        return fut

    def unsubscribe(self, executor_id: str) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        fut = self._m.control.drop_table(self.table_id, {executor_id})
        fut.add_listener(lambda _: self._m.subscriptions.unregister_subscription(self.table_id, executor_id))
        return fut

    def associate(self, executors: Sequence[AllocatedExecutor]) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        ids = {e.id if isinstance(e, AllocatedExecutor) else e for e in executors}
        already = ids & self._m.subscriptions.get_subscribers(self.table_id)
        new_ids = ids - already
        op = self._m.subscriptions.start_subscription_init(self.table_id, new_ids)
        fut = self._m.control.init_table(self.conf, new_ids, self.block_manager.get_ownership_status())

        def done(_):
            for e in ids:
                self.block_manager.register_executor(e)
            self._m.subscriptions.finish_subscription_init(op)
        if self.is_device and new_ids:
            out: ResultFuture = ResultFuture()

            def after(acks):
                def fin(_):
                    done(None)
                    out.on_completed(None)
                self._exchange_peer_map(acks, new_ids).chain(out, fin)
            fut.chain(out, after)
            return out
        fut.add_listener(done)
        return fut

    def unassociate(self, executor_id: str) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        result: ResultFuture = ResultFuture()
        to_sync = self._m.subscriptions.get_subscribers(self.table_id) - {executor_id}

        def after_sync(_):
            def after_drop(_):
                self.block_manager.deregister_executor(executor_id)
                self._m.subscriptions.unregister_subscription(self.table_id, executor_id)
                result.on_completed(None)
            self._m.control.drop_table(self.table_id, {executor_id}).add_listener(after_drop)
        self._m.control.sync_ownership(self.table_id, executor_id, to_sync).add_listener(after_sync)
        return result

    # ---- elasticity ----
    def move_blocks(self, src: str, dst: str, num_blocks: int) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        if dst not in self.block_manager.get_partition_info():
            raise NotAssociatedError(f"{dst} is not associated with {self.table_id}")
        if src == dst:
            return CompletedFuture(MigrationResult(True, "source and destination are the same executor", []))
        blocks = self.block_manager.choose_blocks_to_move(src, num_blocks)
        if not blocks:
            return CompletedFuture(MigrationResult(False, f"There is no block to move in {src}", []))
        with self._cond:   # checkpoints and migrations are mutually exclusive per table
            self._cond.wait_for(lambda: self._ongoing_chkps == 0)
            self._ongoing_migrations += 1
        if self.is_device:
            fut = self._device_move(src, dst, blocks)
        else:
            fut = self._m.migrations.start_migration(self.block_manager, self.table_id, src, dst, blocks)

        def done(_):
            with self._cond:
                self._ongoing_migrations -= 1
                self._cond.notify_all()
        fut.add_listener(done)
        return fut

    def _device_move(self, src: str, dst: str, blocks: List[int]) -> ResultFuture:
        """Live device migration of a LIST of blocks -- no accessor is stopped:
        alloc at the receiver -> ONE bulk-TMA copy launch at the sender (new slabs over NVLink + local snapshots) ->
        every executor flips its pointer-table entries and fences -> the sender forwards what still landed in the
        old slabs (new += old - snapshot) and frees them.  Other blocks / tables / jobs never notice."""
        out: ResultFuture = ResultFuture()
        subs = self._m.subscriptions.get_subscribers(self.table_id) | self.block_manager.get_associator_ids()
        ctl = self._m.control
        stats: Dict[str, Any] = {}

        def release(exc, free_at_dst: bool):
            LOG.error("device migration of %s blocks %s -> %s failed: %s", self.table_id, src, dst, exc)
            for b in blocks:
                self.block_manager.release_block_from_move(b)
            if free_at_dst:
                ctl.fan_out({dst}, {"type": "DeviceMoveAbort", "table_id": self.table_id,
                                    "block_ids": list(blocks)}).add_listener(lambda _: out.on_failed(exc))
            else:
                out.on_failed(exc)

        def step_done(acks):
            stats["forwarded"] = sum(a.get("forwarded", 0) for a in acks)
            res = MigrationResult(True, "device migration completed", list(blocks))
            res.stats = dict(stats)
            out.on_completed(res)

        def step_drain(_):
            ctl.fan_out({src}, {"type": "DeviceMoveDrain", "table_id": self.table_id,
                                "block_ids": list(blocks)}).chain(out, step_done)

        def step_flip(locs, acks):
            stats.update({"copy_ms": acks[0].get("copy_ms"), "bytes": acks[0].get("bytes")})
            for b in blocks:
                self.block_manager.update_owner(b, src, dst)
                self.block_manager.release_block_from_move(b)
                # keep the driver's slab map current for executors that subscribe later
                key, si, off, seg_info = locs[b] if b in locs else locs[str(b)]
                self.device_infos.get(src, {}).get("blocks", {}).pop(b, None)
                di = self.device_infos.setdefault(dst, {"key": dst, "blocks": {}, "handles": {}, "ptr_tbl": None})
                di["blocks"][b] = (key, si, off)
                di["handles"][si] = seg_info
            ctl.fan_out(subs, {"type": "DeviceOwnershipFlip", "table_id": self.table_id, "locs": locs,
                               "new_owner": dst, "old_owner": src}).chain(out, step_drain)

        def step_send(acks):
            locs = acks[0]["locs"]
            ctl.fan_out({src}, {"type": "DeviceMoveSend", "table_id": self.table_id, "locs": locs}).chain(
                out, lambda a: step_flip(locs, a), on_error=lambda e: release(e, True))
        ctl.fan_out({dst}, {"type": "DeviceMoveAlloc", "table_id": self.table_id,
                            "block_ids": list(blocks)}).chain(out, step_send, on_error=lambda e: release(e, False))
        return out

    def get_partition_info(self) -> Dict[str, Set[int]]:
        return self.block_manager.get_partition_info()

    def get_ownership_status(self) -> List[str]:
        return self.block_manager.get_ownership_status()

    def get_owner_id(self, block_id: int) -> str:
        return self.block_manager.get_owner_id(block_id)

    # ---- checkpoint ----
    def checkpoint(self, sampling_ratio: float = 1.0) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        if not (0 < sampling_ratio <= 1):
            raise ValueError("Sample ratio should be a positive value that is less than and equal to 1.")
        with self._cond:
            self._cond.wait_for(lambda: self._ongoing_migrations == 0)
            self._ongoing_chkps += 1
        fut = self._m.chkp_master.checkpoint(self.conf, self.block_manager.get_associator_ids(), sampling_ratio)

        def done(_):
            with self._cond:
                self._ongoing_chkps -= 1
                self._cond.notify_all()
        fut.add_listener(done)
        return fut

    # ---- drop ----
    def drop(self) -> ResultFuture:
        self.sm.check_state("INITIALIZED")
        targets = self.block_manager.get_associator_ids() | self._m.subscriptions.unregister_subscribers(self.table_id)
        targets &= set(self._m.executors)
        fut: ResultFuture = ResultFuture()

        def done(_):
            self.sm.set_state("DROPPED")
            self._m.table_manager.on_drop_table(self.table_id)
            fut.on_completed(None)
        self._m.control.drop_table(self.table_id, targets).add_listener(done)
        return fut


class TableManager:
    """Driver registry of tables: create (fresh / checkpoint / input), get, drop bookkeeping.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/TableManager.java:75-157."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self._tables: Dict[str, AllocatedTable] = {}
        self._lock = threading.Lock()

    def create_table(self, conf: TableConfiguration, executors: Sequence[AllocatedExecutor]) -> ResultFuture:
        with self._lock:
            if conf.table_id in self._tables:
                raise RuntimeError(f"Table {conf.table_id} already exists")
            t = AllocatedTable(self._m)
            self._tables[conf.table_id] = t
        return t.init(conf, executors)

    def create_table_from_chkp(self, chkp_id: str, executors: Sequence[AllocatedExecutor],
                               table_id: Optional[str] = None) -> ResultFuture:
        t = AllocatedTable(self._m)
        fut = t.init_from_chkp(chkp_id, executors, table_id)
        with self._lock:
            if t.table_id in self._tables:
                raise RuntimeError(f"Table {t.table_id} already exists")
            self._tables[t.table_id] = t
        return fut

    def get_allocated_table(self, table_id: str) -> AllocatedTable:
        with self._lock:
            t = self._tables.get(table_id)
        if t is None:
            raise TableNotExistError(table_id)
        return t

    def has_table(self, table_id: str) -> bool:
        with self._lock:
            return table_id in self._tables

    def on_drop_table(self, table_id: str) -> None:
        with self._lock:
            self._tables.pop(table_id, None)

    def table_ids(self) -> List[str]:
        with self._lock:
            return list(self._tables)


# ==========================================================================================
# Metrics (driver)
# ==========================================================================================
class MetricReceiver:
    """SPI receiving executor metric reports.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/api/MetricReceiver.java:26-34."""

    def on_metric_report(self, report: Dict[str, Any]) -> None:   # default: ignore
        pass


class MetricManager:
    """Starts / stops metric collection on executors and forwards reports to the receivers.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/metric/MetricManager.java:46-56."""

    def __init__(self, master: "ETMaster") -> None:
        self._m = master
        self.receivers: List[MetricReceiver] = []

    def start_metric_collection(self, executor_id: str, flush_period_ms: int = 0) -> None:
        self._m.send(executor_id, {"type": "MetricControl", "command": "Start", "flush_period_ms": flush_period_ms})

    def stop_metric_collection(self, executor_id: str) -> None:
        self._m.send(executor_id, {"type": "MetricControl", "command": "Stop"})

    def on_report(self, m: Dict[str, Any]) -> None:
        for r in list(self.receivers):
            try:
                r.on_metric_report(m)
            except Exception:  # noqa: BLE001
                LOG.exception("metric receiver failed")


# ==========================================================================================
# ETMaster
# ==========================================================================================
class ETMaster:
    """Driver entry point of Elastic Table: add executors, create tables, look executors / tables up; owns the
    managers, the message demux and the failure listeners.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/driver/impl/ETMasterImpl.java:44-88,
    driver/impl/MessageHandlerImpl.java:109-141,216-232."""

    def __init__(self, transport: Transport, launcher: Optional[ExecutorLauncher] = None,
                 driver_id: str = "driver", chkp_temp_path: str = "/tmp/chkp",
                 chkp_commit_path: str = "/tmp/chkp/commit") -> None:
        self.transport = transport
        self.driver_id = driver_id
        self.launcher = launcher or InProcessLauncher()
        self.executors: Dict[str, AllocatedExecutor] = {}
        self._pending_exec: Dict[str, ResultFuture] = {}
        self._pending_conf: Dict[str, ExecutorConfiguration] = {}
        self._tasklets: Dict[Tuple[str, str], RunningTasklet] = {}
        self._exec_ids = itertools.count()
        self._lock = threading.RLock()
        self.pool = new_fixed_thread_pool(16, "driver-pool")
        self.subscriptions = SubscriptionManager(self)
        self.migrations = MigrationManager(self)
        self.control = TableControlAgent(self)
        self.chkp_master = ChkpManagerMaster(self, chkp_temp_path, chkp_commit_path)
        self.fallback = FallbackManager(self)
        self.task_units = GlobalTaskUnitScheduler(self)
        self.table_manager = TableManager(self)
        self.metrics = MetricManager(self)
        self.custom_handlers: Dict[str, Callable[[Dict[str, Any]], None]] = {}   # msg type -> handler
        self.centcomm_handlers: Dict[str, Callable[[Dict[str, Any]], None]] = {}
        self.failure_listeners: List[Callable[[str, str], None]] = []
        self.chkp_temp_path, self.chkp_commit_path = chkp_temp_path, chkp_commit_path
        transport.register(driver_id, self.on_message)

    # ---- messaging ----
    def send(self, dst: str, msg: Dict[str, Any]) -> None:
        msg.setdefault("src", self.driver_id)
        self.transport.send(dst, msg)

    def on_message(self, m: Dict[str, Any]) -> None:
        t = m["type"]
        if t in ("TableInitAck", "TableLoadAck", "TableDropAck", "OwnershipSyncAck", "TablePeerMapAck",
                 "TableQuiesceAck", "TableResumeAck", "DeviceMoveAllocAck", "DeviceMoveSendAck", "DeviceMoveDrainAck",
                 "DeviceMoveAbortAck", "DeviceOwnershipFlipAck"):
            self.control.on_ack(m["op_id"], m)
        elif t == "ControlOpFailed":
            LOG.error("control op %s (%s) failed on %s: %s", m["op_id"], m.get("orig_type"), m.get("src"), m["error"])
            self.control.on_failed(m["op_id"], f"{m.get('orig_type')} failed on {m.get('src')}: {m['error']}")
        elif t == "OwnershipMoved":
            self.pool.submit(self.migrations.on_ownership_moved, m["op_id"], m["block_id"])
        elif t == "DataMoved":
            self.pool.submit(self.migrations.on_data_moved, m["op_id"], m["block_id"], m["ownership_together"])
        elif t == "TableAccessReq":
            self.pool.submit(self.fallback.on_table_access_req, m)
        elif t == "OwnershipReq":
            self.pool.submit(self._on_ownership_req, m)
        elif t == "ChkpFailed":
            self.chkp_master.on_chkp_failed(m)
        elif t == "ChkpDone":
            self.chkp_master.on_chkp_done(m["chkp_id"], m["executor_id"], m["block_ids"])
        elif t == "ChkpCommitAck":
            self.chkp_master.on_commit_ack(m["chkp_id"], m["executor_id"])
        elif t == "ChkpLoadDone":
            self.chkp_master.on_load_done(m["op_id"], m.get("num_items"))
        elif t == "TaskUnitWait":
            self.task_units.on_wait(m)
        elif t == "TaskUnitJoin":
            self.task_units.on_join(m["job_id"], m["executor_id"])
        elif t == "TaskUnitLeave":
            self.task_units.on_leave(m["job_id"], m["executor_id"])
        elif t == "TaskletStatus":
            with self._lock:
                rt = self._tasklets.get((m["executor_id"], m["tasklet_id"]))
            if rt is not None:
                rt.on_status(m)
                if m["status"] in ("Done", "Failed"):
                    # the handle (and its result) lives on with whoever holds it; the registries of a long-running
                    # driver only keep live tasklets
                    with self._lock:
                        self._tasklets.pop((m["executor_id"], m["tasklet_id"]), None)
                    ae = self.executors.get(m["executor_id"])
                    if ae is not None:
                        ae.forget_tasklet(m["tasklet_id"])
        elif t == "TaskletCustom":
            with self._lock:
                rt = self._tasklets.get((m["executor_id"], m["tasklet_id"]))
            # `inline` = the sender vouches that the handler never blocks (per-mini-batch sync / progress): run it
            # here, on the inbox thread, and save the pool hop (the inbox loop logs a failing handler)
            if rt is not None and rt.custom_handler is not None:
                if m.get("inline"):
                    rt.custom_handler(m["payload"])
                else:
                    self.pool.submit(rt.custom_handler, m["payload"])
            elif "TaskletCustom" in self.custom_handlers:
                if m.get("inline"):
                    self.custom_handlers["TaskletCustom"](m)
                else:
                    self.pool.submit(self.custom_handlers["TaskletCustom"], m)
        elif t == "MetricReport":
            self.pool.submit(self.metrics.on_report, m)
        elif t == "ExecutorRegister":
            self.on_executor_registered(m["executor_id"])
        elif t == "ExecutorClosed":
            self._on_executor_closed(m["executor_id"])
        elif t == "CentComm":
            h = self.centcomm_handlers.get(m.get("client"))
            if h is not None:
                self.pool.submit(h, m)
        elif t in self.custom_handlers:
            self.pool.submit(self.custom_handlers[t], m)
        else:
            LOG.warning("driver: unhandled message %s", t)

    def _on_ownership_req(self, m: Dict[str, Any]) -> None:
        try:
            table = self.table_manager.get_allocated_table(m["table_id"])
        except TableNotExistError:
            return
        owner = table.block_manager.get_owner_id(m["block_id"])
        try:   # `old == driver_id` marks a reply (not a broadcast)
            self.send(m["src"], {"type": "OwnershipUpdate", "table_id": m["table_id"], "block_id": m["block_id"],
                                 "old": self.driver_id, "new": owner})
        except NetworkError:
            pass

    # ---- executors ----
    def add_executors(self, num: int, conf: Optional[ExecutorConfiguration] = None) -> ResultFuture:
        conf = conf or ExecutorConfiguration()
        agg = AggregateFuture(num)
        for _ in range(num):
            with self._lock:
                eid = f"executor-{next(self._exec_ids)}"
                fut: ResultFuture = ResultFuture()
                self._pending_exec[eid] = fut
                self._pending_conf[eid] = conf
            fut.add_listener(agg.on_partial)
            self.launcher.launch(eid, conf, self)
        return agg

    def on_executor_registered(self, executor_id: str) -> None:
        with self._lock:
            fut = self._pending_exec.pop(executor_id, None)
            conf = self._pending_conf.pop(executor_id, ExecutorConfiguration())
            ae = AllocatedExecutor(self, executor_id, conf)
            self.executors[executor_id] = ae
        if fut is not None:
            fut.on_completed(ae)

    def _on_executor_closed(self, executor_id: str) -> None:
        with self._lock:
            ae = self.executors.pop(executor_id, None)
        self.chkp_master.on_executor_closed(executor_id)
        if ae is not None:
            ae.on_closed()
        self.launcher.terminate(executor_id)

    def on_executor_failed(self, executor_id: str, reason: str = "") -> None:
        """EvaluatorFailedHandler (driver/impl/EvaluatorFailedHandler.java:39-42): the executor is
        unregistered; tasklets it was running are reported as Failed so their jobs fail instead of
        hanging (the reference throws in the driver, TODO #677)."""
        LOG.error("executor %s failed: %s", executor_id, reason)
        with self._lock:
            ae = self.executors.pop(executor_id, None)
            rts = [rt for (eid, _), rt in self._tasklets.items() if eid == executor_id and not rt.is_done()]
        for rt in rts:
            try:
                rt.sm.set_state("FAILED")
            except Exception:  # noqa: BLE001
                pass
            rt._result.on_completed(TaskletResult(rt.tasklet_id, executor_id, "Failed", None,
                                                  f"executor {executor_id} failed: {reason}"))
        self.chkp_master.on_executor_closed(executor_id)
        if ae is not None:
            ae.on_closed()
        try:
            self.launcher.terminate(executor_id)        # reclaim whatever is left of it
        except Exception:  # noqa: BLE001
            LOG.exception("terminate of failed executor %s", executor_id)
        for cb in list(self.failure_listeners):
            try:
                cb(executor_id, reason)
            except Exception:  # noqa: BLE001
                LOG.exception("failure listener failed")

    def get_executor(self, executor_id: str) -> AllocatedExecutor:
        with self._lock:
            ae = self.executors.get(executor_id)
        if ae is None:
            raise ExecutorNotExistError(executor_id)
        return ae

    def register_tasklet(self, rt: RunningTasklet) -> None:
        with self._lock:
            self._tasklets[(rt.executor_id, rt.tasklet_id)] = rt

    # ---- tables ----
    def create_table(self, conf: TableConfiguration, executors: Sequence[AllocatedExecutor]) -> ResultFuture:
        return self.table_manager.create_table(conf, executors)

    def create_table_from_chkp(self, chkp_id: str, executors: Sequence[AllocatedExecutor],
                               table_id: Optional[str] = None) -> ResultFuture:
        return self.table_manager.create_table_from_chkp(chkp_id, executors, table_id)

    def get_table(self, table_id: str) -> AllocatedTable:
        return self.table_manager.get_allocated_table(table_id)

    def has_table(self, table_id: str) -> bool:
        return self.table_manager.has_table(table_id)

    def close(self) -> None:
        for ae in list(self.executors.values()):
            try:
                ae.close().get(timeout=10)
            except Exception:  # noqa: BLE001
                pass
        self.transport.unregister(self.driver_id)
        self.pool.shutdown(wait=False)
