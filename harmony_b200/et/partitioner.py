"""key -> block partitioners.

Reference: HashBasedBlockPartitioner.java:40-55 (|murmur(encodedKey)| / (MAX_INT/numBlocks)),
OrderingBasedBlockPartitioner.java:38-58 (key / (Long.MAX/numBlocks), contiguous key range per
block), KeySpacePartitioner.java:28-44.

B200 design: dense tables address a bounded integer key space [0, capacity) so that a block is a
contiguous HBM slab; `ordered` = range partitioning (block b owns [b*R, (b+1)*R)), `hashed` =
strided partitioning (key % num_blocks -- a perfect hash of a dense key space).  The same two
mappings are evaluated on the device by hb::row_ptr (ops/csrc/hb_common.cuh).  Generic (dict)
tables accept arbitrary hashable keys through murmur3.
"""
from __future__ import annotations

import pickle
from typing import Any, Sequence, Tuple

import numpy as np

from ..utils.misc import murmur3_32

PART_ORDERED, PART_HASHED = 0, 1
_INT_MAX = 2 ** 31 - 1
_LONG_MAX = 2 ** 63 - 1


class BlockPartitioner:
    """key -> block id: range partitioning for ordered tables, hashing otherwise; key ranges of blocks.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/OrderingBasedBlockPartitioner.java,
    HashBasedBlockPartitioner.java."""

    num_blocks: int

    def get_block_id(self, key: Any) -> int:
        raise NotImplementedError

    def get_block_ids(self, keys: Sequence[Any]) -> np.ndarray:
        return np.fromiter((self.get_block_id(k) for k in keys), dtype=np.int64, count=len(keys))


class DenseOrderedPartitioner(BlockPartitioner):
    """block b owns keys [b*rows_per_block, (b+1)*rows_per_block)."""
    mode = PART_ORDERED

    def __init__(self, num_blocks: int, capacity: int) -> None:
        self.num_blocks = num_blocks
        self.capacity = capacity
        self.rows_per_block = max(1, -(-capacity // num_blocks))

    def get_block_id(self, key):
        return int(key) // self.rows_per_block

    def get_block_ids(self, keys):
        return np.asarray(keys, dtype=np.int64) // self.rows_per_block

    def slot(self, keys):
        return np.asarray(keys, dtype=np.int64) % self.rows_per_block

    def key_range(self, block_id: int) -> Tuple[int, int]:
        lo = block_id * self.rows_per_block
        return lo, min(self.capacity, lo + self.rows_per_block) - 1

    def keys_of_block(self, block_id: int) -> np.ndarray:
        lo, hi = self.key_range(block_id)
        return np.arange(lo, hi + 1, dtype=np.int64)


class DenseHashedPartitioner(BlockPartitioner):
    """block = key % num_blocks, slot = key // num_blocks (perfect hash of a dense key space)."""
    mode = PART_HASHED

    def __init__(self, num_blocks: int, capacity: int) -> None:
        self.num_blocks = num_blocks
        self.capacity = capacity
        self.rows_per_block = max(1, -(-capacity // num_blocks))

    def get_block_id(self, key):
        return int(key) % self.num_blocks

    def get_block_ids(self, keys):
        return np.asarray(keys, dtype=np.int64) % self.num_blocks

    def slot(self, keys):
        return np.asarray(keys, dtype=np.int64) // self.num_blocks

    def keys_of_block(self, block_id: int) -> np.ndarray:
        return np.arange(block_id, self.capacity, self.num_blocks, dtype=np.int64)


class OrderingBasedBlockPartitioner(BlockPartitioner):
    """Generic ordered tables over the whole non-negative long space (reference semantics)."""

    def __init__(self, num_blocks: int) -> None:
        self.num_blocks = num_blocks
        self.block_size = _LONG_MAX // num_blocks

    def get_block_id(self, key):
        return min(int(key) // self.block_size, self.num_blocks - 1)

    def key_range(self, block_id: int) -> Tuple[int, int]:
        lo = block_id * self.block_size
        hi = _LONG_MAX if block_id == self.num_blocks - 1 else lo + self.block_size - 1
        return lo, hi


def encode_key(key: Any) -> bytes:
    if isinstance(key, (int, np.integer)):
        return int(key).to_bytes(8, "big", signed=True)
    if isinstance(key, str):
        return key.encode("utf-8")
    if isinstance(key, bytes):
        return key
    return pickle.dumps(key, protocol=4)


class HashBasedBlockPartitioner(BlockPartitioner):
    """Generic hashed tables: |murmur3(encoded key)| / (INT_MAX / num_blocks)."""

    def __init__(self, num_blocks: int) -> None:
        self.num_blocks = num_blocks
        self.key_space_size = max(1, _INT_MAX // num_blocks)

    def get_block_id(self, key):
        h = murmur3_32(encode_key(key))
        if h >= 2 ** 31:
            h -= 2 ** 32
        return min(abs(h) // self.key_space_size, self.num_blocks - 1)
