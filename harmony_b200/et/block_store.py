"""Local shard of a table: blocks, the block store and the Tablet view.

Reference: services/et/.../evaluator/impl/BlockImpl.java:65-76 (update = compute(k, v ->
updateFunction.updateValue(k, v ?? initValue(k), u))), BlockStore.java:55-117,
TabletImpl.java:96-190, evaluator/api/Tablet.java:34-128.

Two host block kinds share one interface:
  * DictBlock  -- arbitrary hashable keys -> python/numpy values (input tables, GBT, Pregel, LDA
                  local model ...), guarded by one lock per block.
  * DenseBlock -- bounded integer key space, numpy slab [rows_per_block, row_dim] + presence
                  mask; multi-key ops are vectorised.  This is the host twin of the HBM slab
                  used by the device backend (et/device_table.py).
"""
from __future__ import annotations

import threading
from typing import Any, Dict, Iterator, List, Optional, Tuple

import numpy as np

from .update_function import UpdateFunction


class BlockNotExistsError(KeyError):
    """The block is not (or no longer) stored on this executor."""

    pass


class Block:
    """One block of a table: per-key ops + bulk export / import for migration and checkpoint.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/BlockImpl.java:65-76,
    BlockStore.java:55-117."""

    block_id: int

    def put(self, key, value): raise NotImplementedError
    def put_if_absent(self, key, value): raise NotImplementedError
    def get(self, key): raise NotImplementedError
    def get_or_init(self, key): raise NotImplementedError
    def remove(self, key): raise NotImplementedError
    def update(self, key, delta): raise NotImplementedError
    def items(self) -> Iterator[Tuple[Any, Any]]: raise NotImplementedError
    def num_items(self) -> int: raise NotImplementedError
    def export_items(self) -> Tuple[Any, Any]: raise NotImplementedError
    def import_items(self, keys, values) -> None: raise NotImplementedError

    # generic multi-key fallbacks
    def multi_put(self, keys, values):
        for k, v in zip(keys, values):
            self.put(k, v)

    def multi_get(self, keys):
        return [self.get(k) for k in keys]

    def multi_get_or_init(self, keys):
        return [self.get_or_init(k) for k in keys]

    def multi_update(self, keys, deltas):
        return [self.update(k, d) for k, d in zip(keys, deltas)]


class BlockSnapshot(list):
    """The (key, value) items of one block at one version: a list shared by every reader until the block changes
    (read-only by convention).  Carries what batch caches key on, so a trainer need not walk the items."""

    __slots__ = ("block_id", "version", "_min_key")

    def __init__(self, items, block_id: int, version: int) -> None:
        super().__init__(items)
        self.block_id, self.version, self._min_key = block_id, version, None

    def min_key(self):
        if self._min_key is None and len(self):
            self._min_key = min(k for k, _ in self)
        return self._min_key


class DictBlock(Block):
    """Generic key-value block (hash map + update function) for non-dense tables."""

    def __init__(self, block_id: int, update_fn: UpdateFunction) -> None:
        self.block_id = block_id
        self._fn = update_fn
        self._map: Dict[Any, Any] = {}
        self._lock = threading.RLock()
        self._version = 0                      # bumped by every mutation: validates the shared snapshot
        self._snap: Optional[BlockSnapshot] = None

    def put(self, key, value):
        with self._lock:
            old = self._map.get(key)
            self._map[key] = value
            self._version += 1
            return old

    def put_if_absent(self, key, value):
        with self._lock:
            if key in self._map:
                return self._map[key]
            self._map[key] = value
            self._version += 1
            return None

    def get(self, key):
        with self._lock:
            return self._map.get(key)

    def get_or_init(self, key):
        with self._lock:
            if key not in self._map:
                self._map[key] = self._fn.init_value(key)
                self._version += 1
            return self._map[key]

    def remove(self, key):
        with self._lock:
            self._version += 1
            return self._map.pop(key, None)

    def update(self, key, delta):
        with self._lock:
            old = self._map[key] if key in self._map else self._fn.init_value(key)
            new = self._fn.update_value(key, old, delta)
            self._map[key] = new
            self._version += 1
            return new

    def items(self):
        with self._lock:
            return iter(list(self._map.items()))

    def items_snapshot(self) -> BlockSnapshot:
        """The items as a list that is rebuilt only after a mutation: the training loop reads every input block once
        per epoch and the data never changes, so epochs after the first copy nothing."""
        with self._lock:
            if self._snap is None or self._snap.version != self._version:
                self._snap = BlockSnapshot(self._map.items(), self.block_id, self._version)
            return self._snap

    def num_items(self):
        with self._lock:
            return len(self._map)

    def export_items(self):
        with self._lock:
            return list(self._map.keys()), list(self._map.values())

    def import_items(self, keys, values):
        with self._lock:
            for k, v in zip(keys, values):
                self._map[k] = v
            self._version += 1

    def clear(self):
        with self._lock:
            self._map.clear()
            self._version += 1
            self._snap = None


class DenseBlock(Block):
    """numpy slab; `slot_of(keys)` maps global keys to local row indices."""

    def __init__(self, block_id: int, update_fn: UpdateFunction, partitioner, row_dim: int, dtype) -> None:
        self.block_id = block_id
        self._fn = update_fn
        self._part = partitioner
        self.row_dim = row_dim
        self.dtype = np.dtype(dtype)
        self.rows = np.zeros((partitioner.rows_per_block, row_dim), dtype=self.dtype)
        self.present = np.zeros(partitioner.rows_per_block, dtype=bool)
        self._lock = threading.RLock()

    def _slots(self, keys) -> np.ndarray:
        return self._part.slot(np.asarray(keys, dtype=np.int64))

    def _ensure(self, keys: np.ndarray, slots: np.ndarray) -> None:
        missing = ~self.present[slots]
        if missing.any():
            mk, ms = keys[missing], slots[missing]
            # de-duplicate so that a key is initialised once
            _, first = np.unique(ms, return_index=True)
            self.rows[ms[first]] = self._fn.init_rows(mk[first], self.row_dim, self.dtype)
            self.present[ms[first]] = True

    # ---- single key ----
    def put(self, key, value):
        with self._lock:
            s = int(self._slots([key])[0])
            old = self.rows[s].copy() if self.present[s] else None
            self.rows[s] = value
            self.present[s] = True
            return old

    def put_if_absent(self, key, value):
        with self._lock:
            s = int(self._slots([key])[0])
            if self.present[s]:
                return self.rows[s].copy()
            self.rows[s] = value
            self.present[s] = True
            return None

    def get(self, key):
        with self._lock:
            s = int(self._slots([key])[0])
            return self.rows[s].copy() if self.present[s] else None

    def get_or_init(self, key):
        return self.multi_get_or_init(np.asarray([key], dtype=np.int64))[0]

    def remove(self, key):
        with self._lock:
            s = int(self._slots([key])[0])
            if not self.present[s]:
                return None
            old = self.rows[s].copy()
            self.present[s] = False
            self.rows[s] = 0
            return old

    def update(self, key, delta):
        return self.multi_update(np.asarray([key], dtype=np.int64), np.asarray(delta)[None, :])[0]

    # ---- multi key (vectorised) ----
    def multi_put(self, keys, values):
        with self._lock:
            s = self._slots(keys)
            self.rows[s] = np.asarray(values, dtype=self.dtype)
            self.present[s] = True

    def multi_get(self, keys):
        with self._lock:
            s = self._slots(keys)
            out = self.rows[s].copy()
            return [out[i] if self.present[s[i]] else None for i in range(len(s))]

    def multi_get_or_init(self, keys):
        with self._lock:
            keys = np.asarray(keys, dtype=np.int64)
            s = self._slots(keys)
            self._ensure(keys, s)
            return self.rows[s].copy()

    def multi_update(self, keys, deltas):
        with self._lock:
            keys = np.asarray(keys, dtype=np.int64)
            s = self._slots(keys)
            self._ensure(keys, s)
            deltas = np.asarray(deltas)
            uniq, inv = np.unique(s, return_inverse=True)
            if len(uniq) == len(s):
                self.rows[s] = self._fn.update_rows(self.rows[s], deltas)
            else:
                # duplicated keys in one request: apply sequentially per duplicate rank
                order = np.argsort(inv, kind="stable")
                rank = np.zeros(len(s), dtype=np.int64)
                sorted_inv = inv[order]
                starts = np.r_[0, np.flatnonzero(np.diff(sorted_inv)) + 1]
                rank[order] = np.arange(len(s)) - np.repeat(starts, np.diff(np.r_[starts, len(s)]))
                for r in range(int(rank.max()) + 1):
                    m = rank == r
                    self.rows[s[m]] = self._fn.update_rows(self.rows[s[m]], deltas[m])
            return self.rows[s].copy()

    def items(self):
        with self._lock:
            keys = self._part.keys_of_block(self.block_id)
            slots = self._slots(keys)
            m = self.present[slots]
            return iter([(int(k), self.rows[s].copy()) for k, s in zip(keys[m], slots[m])])

    def num_items(self):
        with self._lock:
            return int(self.present.sum())

    def export_items(self):
        with self._lock:
            keys = self._part.keys_of_block(self.block_id)
            slots = self._slots(keys)
            m = self.present[slots]
            return keys[m].copy(), self.rows[slots[m]].copy()

    def import_items(self, keys, values):
        if len(keys):
            self.multi_put(np.asarray(keys, dtype=np.int64), values)

    def clear(self):
        with self._lock:
            self.rows[:] = 0
            self.present[:] = False


class BlockStore:
    """block id -> Block for the blocks this executor currently owns."""

    def __init__(self, block_factory) -> None:
        self._factory = block_factory
        self._blocks: Dict[int, Block] = {}
        self._lock = threading.RLock()

    def create_empty_block(self, block_id: int) -> Block:
        with self._lock:
            if block_id in self._blocks:
                raise RuntimeError(f"Block with id {block_id} already exists")
            b = self._factory(block_id)
            self._blocks[block_id] = b
            return b

    def put_block(self, block_id: int, keys, values) -> None:
        b = self.create_empty_block(block_id)
        b.import_items(keys, values)

    def remove_block(self, block_id: int) -> None:
        with self._lock:
            if self._blocks.pop(block_id, None) is None:
                raise BlockNotExistsError(block_id)

    def get(self, block_id: int) -> Block:
        with self._lock:
            b = self._blocks.get(block_id)
        if b is None:
            raise BlockNotExistsError(block_id)
        return b

    def try_get(self, block_id: int) -> Optional[Block]:
        with self._lock:
            return self._blocks.get(block_id)

    def block_ids(self) -> List[int]:
        with self._lock:
            return sorted(self._blocks)

    def num_blocks(self) -> int:
        with self._lock:
            return len(self._blocks)

    def __iter__(self) -> Iterator[Block]:
        with self._lock:
            return iter(list(self._blocks.values()))


class Tablet:
    """Read view over the local blocks (reference Tablet API: Tablet.java:34-128)."""

    def __init__(self, store: BlockStore) -> None:
        self._store = store

    def get_data_map(self) -> Dict[Any, Any]:
        out: Dict[Any, Any] = {}
        for b in self._store:
            out.update(dict(b.items()))
        return out

    def get_num_blocks(self) -> int:
        return self._store.num_blocks()

    def get_block_ids(self) -> List[int]:
        return self._store.block_ids()

    def get_num_data_items(self) -> int:
        return sum(b.num_items() for b in self._store)

    def get_block(self, block_id: int) -> Block:
        return self._store.get(block_id)

    def get_block_iterator(self) -> Iterator[Block]:
        return iter(self._store)

    def get_data_iterator(self) -> Iterator[Tuple[Any, Any]]:
        for b in self._store:
            yield from b.items()

    def get_data_in_block(self, block_id: int) -> List[Tuple[Any, Any]]:
        b = self._store.try_get(block_id)
        if b is None:
            return []
        snap = getattr(b, "items_snapshot", None)
        return snap() if snap is not None else list(b.items())
