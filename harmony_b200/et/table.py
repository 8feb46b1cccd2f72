"""Table API (evaluator side) + the per-executor table registry.

Reference: evaluator/api/Table.java:46-221 (17 ops, all returning futures),
evaluator/impl/TableImpl.java:109-143,284-326,366-408,420-448,460-501 (routing: encode key ->
block -> ownership read lock -> local tablet or remote sender; multi-key ops grouped by block,
one message per block; UPDATE always goes through the comm queue even when local),
evaluator/impl/Tables.java:79-193 + TableComponents.java:28-123.
"""
from __future__ import annotations

import logging
import threading
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from ..utils.concurrent import ResultFuture
from .block_store import BlockStore, DenseBlock, DictBlock, Tablet
from .config import TableConfiguration
from .exceptions import TableNotExistError
from .ownership import OwnershipCache
from .partitioner import (DenseHashedPartitioner, DenseOrderedPartitioner, HashBasedBlockPartitioner,
                          OrderingBasedBlockPartitioner)
from .remote_access import (GET, GET_OR_INIT, PUT, PUT_IF_ABSENT, REMOVE, UPDATE, MultiKeyDataOpResult, RemoteDataOp, SingleKeyDataOpResult)
from .update_function import make_update_function

LOG = logging.getLogger(__name__)


def make_partitioner(conf: TableConfiguration):
    if conf.dense:
        cls = DenseOrderedPartitioner if conf.is_ordered else DenseHashedPartitioner
        return cls(conf.num_total_blocks, conf.capacity)
    return OrderingBasedBlockPartitioner(conf.num_total_blocks) if conf.is_ordered \
        else HashBasedBlockPartitioner(conf.num_total_blocks)


class TableComponents:
    """Everything an executor holds for one table."""

    def __init__(self, conf: TableConfiguration, executor) -> None:
        self.conf = conf
        self.update_fn = make_update_function(conf.update_function)
        if conf.dense:
            self.update_fn.params.setdefault("row_dim", conf.row_dim)
            self.update_fn.params.setdefault("dtype", conf.dtype)
        self.partitioner = make_partitioner(conf)
        self.ownership = OwnershipCache(conf.table_id, conf.num_total_blocks, executor.executor_id,
                                        executor.driver_id, sync_ack=executor.send_ownership_sync_ack)
        if conf.dense:
            factory = lambda b: DenseBlock(b, self.update_fn, self.partitioner, conf.row_dim, conf.dtype)  # noqa: E731
        else:
            factory = lambda b: DictBlock(b, self.update_fn)  # noqa: E731
        self.block_store = BlockStore(factory)
        self.tablet = Tablet(self.block_store)
        self.table: Optional["TableImpl"] = None


class TableImpl:
    """The user-facing table handle of one executor."""

    def __init__(self, comps: TableComponents, executor) -> None:
        self._c = comps
        self._ex = executor
        self.table_id = comps.conf.table_id
        self.conf = comps.conf

    # ---- helpers ----
    def _block_of(self, key) -> int:
        return self._c.partitioner.get_block_id(key)

    def _single(self, op_type: str, key, value, reply_required: bool) -> ResultFuture:
        fut = SingleKeyDataOpResult()
        block_id = self._block_of(key)

        def cb(values, ok):
            if ok:
                fut.on_completed(values[0] if values else None)
            else:
                fut.on_failed(RuntimeError(f"{op_type} failed for key {key!r}"))

        op = RemoteDataOp(self._ex.sender.next_op_id(), self._ex.executor_id, self.table_id, op_type, block_id,
                          [key], None if value is None else [value], reply_required, True,
                          callback=cb if reply_required else None)
        owner, unlock = self._c.ownership.resolve_executor_with_lock(block_id)
        try:
            if owner is None:
                self._ex.handler.execute_local(op, self._c)      # inline, under the ownership read lock
            else:
                self._ex.sender.send_op(op, owner)
        finally:
            unlock()
        if not reply_required:
            fut.on_completed(None)
        return fut

    def _multi(self, op_type: str, keys: Sequence[Any], values: Optional[Sequence[Any]],
               reply_required: bool) -> ResultFuture:
        n = len(keys)
        dense = self.conf.dense
        if dense:
            keys_arr = np.asarray(keys, dtype=np.int64)
            blocks = self._c.partitioner.get_block_ids(keys_arr)
        else:
            keys_arr = list(keys)
            blocks = np.fromiter((self._block_of(k) for k in keys_arr), dtype=np.int64, count=n)
        order = np.argsort(blocks, kind="stable")
        sorted_blocks = blocks[order]
        bounds = np.flatnonzero(np.diff(sorted_blocks)) + 1 if n else np.array([], dtype=np.int64)
        groups = np.split(order, bounds) if n else []
        fut = MultiKeyDataOpResult(n, len(groups), dense and op_type in (GET_OR_INIT, UPDATE))
        if values is not None and dense:
            values = np.asarray(values)
        by_owner: Dict[str, List[RemoteDataOp]] = {}
        unlocks = []
        try:
            for pos in groups:
                block_id = int(blocks[pos[0]])
                if dense:
                    k = keys_arr[pos]
                    v = None if values is None else values[pos]
                else:
                    k = [keys_arr[i] for i in pos]
                    v = None if values is None else [values[i] for i in pos]

                def cb(vals, ok, _pos=pos):
                    fut.on_block_completed(_pos, vals, ok)

                op = RemoteDataOp(self._ex.sender.next_op_id(), self._ex.executor_id, self.table_id, op_type, block_id,
                                  k, v, reply_required, False, callback=cb if reply_required else None)
                owner, unlock = self._c.ownership.resolve_executor_with_lock(block_id)
                if owner is None:
                    try:
                        self._ex.handler.execute_local(op, self._c)  # inline, under the ownership read lock
                    finally:
                        unlock()
                else:
                    # remote block ops are grouped per owner; the blocks' read locks are held until the message
                    # is on its way (same ordering guarantee vs ownership updates as the per-block send)
                    unlocks.append(unlock)
                    by_owner.setdefault(owner, []).append(op)
                if not reply_required:
                    fut.on_block_completed(pos, None, True)
            for owner, ops in by_owner.items():
                self._ex.sender.send_ops(ops, owner)
        finally:
            for u in unlocks:
                u()
        return fut

    # ---- the 17 ops ----
    def put(self, key, value) -> ResultFuture:
        return self._single(PUT, key, value, True)

    def put_no_reply(self, key, value) -> None:
        self._single(PUT, key, value, False)

    def put_if_absent(self, key, value) -> ResultFuture:
        return self._single(PUT_IF_ABSENT, key, value, True)

    def put_if_absent_no_reply(self, key, value) -> None:
        self._single(PUT_IF_ABSENT, key, value, False)

    def multi_put(self, keys, values) -> ResultFuture:
        return self._multi(PUT, keys, values, True)

    def get(self, key) -> ResultFuture:
        return self._single(GET, key, None, True)

    def multi_get(self, keys) -> ResultFuture:
        return self._multi(GET, keys, None, True)

    def get_or_init(self, key) -> ResultFuture:
        return self._single(GET_OR_INIT, key, None, True)

    def multi_get_or_init(self, keys) -> ResultFuture:
        return self._multi(GET_OR_INIT, keys, None, True)

    def update(self, key, delta) -> ResultFuture:
        return self._single(UPDATE, key, delta, True)

    def update_no_reply(self, key, delta) -> None:
        self._single(UPDATE, key, delta, False)

    def multi_update(self, keys, deltas) -> ResultFuture:
        return self._multi(UPDATE, keys, deltas, True)

    def multi_update_no_reply(self, keys, deltas) -> None:
        self._multi(UPDATE, keys, deltas, False)

    def remove(self, key) -> ResultFuture:
        return self._single(REMOVE, key, None, True)

    def remove_no_reply(self, key) -> None:
        self._single(REMOVE, key, None, False)

    def get_local_tablet(self) -> Tablet:
        return self._c.tablet

    # ---- extras ----
    @property
    def components(self) -> TableComponents:
        return self._c

    def get_ownership(self) -> OwnershipCache:
        return self._c.ownership


class Tables:
    """table id -> TableComponents registry of one executor."""

    def __init__(self, executor) -> None:
        self._ex = executor
        self._tables: Dict[str, TableComponents] = {}
        self._lock = threading.RLock()

    def init_table(self, conf: TableConfiguration, block_owners: List[str]):
        with self._lock:
            if conf.table_id in self._tables:
                raise RuntimeError(f"Table {conf.table_id} already exists on {self._ex.executor_id}")
            comps = self._ex.make_table_components(conf)
            comps.ownership.init(block_owners)
            for b, owner in enumerate(block_owners):
                if owner == self._ex.executor_id:
                    comps.block_store.create_empty_block(b)
            comps.table = self._ex.make_table(comps)
            self._tables[conf.table_id] = comps
            return comps.table

    def get_table(self, table_id: str):
        comps = self.try_get_components(table_id)
        if comps is None:
            raise TableNotExistError(f"Table {table_id} does not exist on {self._ex.executor_id}")
        return comps.table

    def get_components(self, table_id: str) -> TableComponents:
        comps = self.try_get_components(table_id)
        if comps is None:
            raise TableNotExistError(table_id)
        return comps

    def try_get_components(self, table_id: str) -> Optional[TableComponents]:
        with self._lock:
            return self._tables.get(table_id)

    def remove(self, table_id: str) -> None:
        with self._lock:
            comps = self._tables.pop(table_id, None)
        if comps is None:
            raise TableNotExistError(table_id)
        closer = getattr(comps, "close", None)
        if closer:
            closer()

    def get_table_to_num_blocks(self) -> Dict[str, int]:
        with self._lock:
            return {tid: c.block_store.num_blocks() for tid, c in self._tables.items()}

    def table_ids(self) -> List[str]:
        with self._lock:
            return list(self._tables)


class TableAccessor:
    """What tasklets use to reach tables (evaluator/api/TableAccessor.java:28-37)."""

    def __init__(self, tables: Tables) -> None:
        self._tables = tables

    def get_table(self, table_id: str):
        return self._tables.get_table(table_id)
