"""OwnershipCache: per-table block -> owner array with per-block RW locks and migration latches.

Behaviour mirrors services/et/.../evaluator/impl/OwnershipCache.java:123-153 (resolve with
read lock, waiting for an incoming block), :195-244 (update under the write lock, arm the latch
when this executor becomes the owner, per-executor block counters that acknowledge an
*unassociation sync* when a leaving executor owns nothing any more) and :285-318 (sync /
allowAccess).  On the device backend the same array is mirrored into HBM as the block pointer
table read by the kernels (et/device_table.py).
"""
from __future__ import annotations

import threading
from typing import Callable, Dict, List, Optional, Tuple

from ..utils.concurrent import CountDownLatch, ReadWriteLock


class OwnershipCache:
    """Replicated block -> owner array with a fair read/write lock per block and incoming-block latches; ops
    resolve the owner under the read lock, ownership updates take the write lock.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/OwnershipCache.java:123-153,195-244,285-318."""

    def __init__(self, table_id: str, num_total_blocks: int, local_executor_id: str, driver_id: str = "driver",
                 sync_ack: Optional[Callable[[int, str, str], None]] = None) -> None:
        self.table_id = table_id
        self.local_executor_id = local_executor_id
        self.driver_id = driver_id
        self._owners: List[Optional[str]] = [None] * num_total_blocks
        self._locks = [ReadWriteLock() for _ in range(num_total_blocks)]
        self._incoming: Dict[int, CountDownLatch] = {}
        self._num_blocks: Dict[str, int] = {}
        self._ongoing_syncs: Dict[str, int] = {}
        self._meta_lock = threading.Lock()
        self._sync_ack = sync_ack
        self._listeners: List[Callable[[int, Optional[str], str], None]] = []

    # ---- init / queries ----
    def init(self, block_owners: List[str]) -> None:
        with self._meta_lock:
            for b, owner in enumerate(block_owners):
                self._owners[b] = owner
                self._num_blocks[owner] = self._num_blocks.get(owner, 0) + 1

    def add_listener(self, cb: Callable[[int, Optional[str], str], None]) -> None:
        self._listeners.append(cb)

    @property
    def num_total_blocks(self) -> int:
        return len(self._owners)

    def owners(self) -> List[Optional[str]]:
        return list(self._owners)

    def resolve_executor(self, block_id: int) -> Optional[str]:
        """None => local, else the remote owner id."""
        owner = self._owners[block_id]
        return None if owner == self.local_executor_id else owner

    def resolve_executor_with_lock(self, block_id: int) -> Tuple[Optional[str], Callable[[], None]]:
        """Takes the block's read lock (caller must invoke the returned unlock) and waits until
        an in-flight migration INTO this executor has delivered the data."""
        lock = self._locks[block_id]
        lock.acquire_read()
        latch = self._incoming.get(block_id)
        if latch is not None:
            latch.await_()
        owner = self._owners[block_id]
        return (None if owner == self.local_executor_id else owner), lock.release_read

    def get_current_local_block_ids(self) -> List[int]:
        return [b for b, o in enumerate(self._owners) if o == self.local_executor_id]

    # ---- updates ----
    def update(self, block_id: int, old_owner_id: str, new_owner_id: str) -> None:
        lock = self._locks[block_id]
        lock.acquire_write()
        try:
            local_old = self._owners[block_id]
            self._owners[block_id] = new_owner_id
            if new_owner_id == self.local_executor_id and local_old != self.local_executor_id:
                self._incoming[block_id] = CountDownLatch(1)
        finally:
            lock.release_write()
        sync_op = None
        with self._meta_lock:
            if local_old is not None:
                rem = self._num_blocks.get(local_old, 0) - 1
                if rem <= 0:
                    self._num_blocks.pop(local_old, None)
                    sync_op = self._ongoing_syncs.pop(local_old, None)
                else:
                    self._num_blocks[local_old] = rem
            self._num_blocks[new_owner_id] = self._num_blocks.get(new_owner_id, 0) + 1
        for cb in self._listeners:
            cb(block_id, local_old, new_owner_id)
        if sync_op is not None:
            self._complete_sync(sync_op, local_old)

    def hold_write_lock(self, block_id: int) -> Callable[[], None]:
        lock = self._locks[block_id]
        lock.acquire_write()
        return lock.release_write

    def allow_access_to_block(self, block_id: int) -> None:
        latch = self._incoming.pop(block_id, None)
        if latch is None:
            raise RuntimeError(f"Block {block_id} is not in migrating state")
        latch.count_down()

    def is_incoming(self, block_id: int) -> bool:
        return block_id in self._incoming

    # ---- unassociation sync ----
    def _complete_sync(self, op_id: int, deleted_executor_id: str) -> None:
        if self._sync_ack is not None:
            self._sync_ack(op_id, self.table_id, deleted_executor_id)

    def sync_unassociation(self, op_id: int, executor_id: str) -> None:
        with self._meta_lock:
            if self._num_blocks.get(executor_id, 0) > 0:
                self._ongoing_syncs[executor_id] = op_id
                return
        self._complete_sync(op_id, executor_id)

    def complete_all_ongoing_sync(self) -> None:
        with self._meta_lock:
            pending = dict(self._ongoing_syncs)
            self._ongoing_syncs.clear()
        for ex, op in pending.items():
            self._complete_sync(op, ex)
