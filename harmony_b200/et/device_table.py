"""Device backend of a dense Elastic Table: block slabs in HBM, shared by all GPUs.

Every block is a slab [rows_per_block, row_stride] in the OWNER GPU's HBM, allocated from a
PeerArena whose segments are CUDA-IPC mapped by all executors.  The replicated ownership array
of the reference (OwnershipCache.java) becomes `block_ptrs`: an int64[num_blocks] tensor in every
GPU's memory holding, for each block, the virtual address of the slab *as mapped in that
process*.  Kernels (ops/csrc/table_ops.cu, mlr_ops.cu ...) resolve key -> (block, slot) -> address
and load / reduce straight into the owner's memory over NVLink.

Table semantics kept from the reference:
  * block -> owner map, initial placement round-robin `block % num_executors` (BlockManager.java:76-88)
  * hashed / ordered partitioning (partitioner.py)
  * update function applied at the owner (fused into the PUSH kernel)
  * lazy initValue becomes eager, key-deterministic initialisation at table creation
  * block migration = bulk slab copy on a side stream + pointer flip on every GPU
    (MigrationExecutor.java:338-414), checkpoint = per-block files (checkpoint.py)
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..parallel.peer import PeerArena, PeerGroup
from .config import TableConfiguration
from .partitioner import PART_HASHED, PART_ORDERED
from .update_function import DEV_ADAGRAD, make_update_function

_CHECK_KEYS = bool(os.environ.get("HB_CHECK_KEYS"))


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class DeviceTable:
    """SPMD handle: every rank of `group` constructs it collectively with the same conf."""

    def __init__(self, conf: TableConfiguration, arena: PeerArena, group: Optional[PeerGroup] = None,
                 block_owners: Optional[Sequence[Any]] = None, init: bool = True, collective: bool = True) -> None:
        """collective=True : SPMD construction (every rank of `group` calls it; handles are exchanged
        with all_gather).  collective=False: job-server construction -- the caller ships
        `local_info()` to the driver and later feeds every peer's info to `apply_infos()`."""
        from ..ops import native
        assert conf.dense, "device tables are dense slab tables"
        self._n = native()
        self.conf = conf
        self.arena = arena
        self.group = group or arena.group
        self.self_key = arena.self_key
        self.rank, self.world = self.group.rank, self.group.world
        self.device = torch.device("cuda", arena.device)
        self.update_fn = make_update_function(conf.update_function)
        if self.update_fn.params.get("coefficient") not in (None, 1):
            raise ValueError("the device ADD update has no coefficient; scale the deltas or use a host table")
        self.torch_dtype = torch.float32 if conf.dtype == "float32" else torch.int32
        self.num_blocks = conf.num_total_blocks
        self.rows_per_block = max(1, -(-conf.capacity // self.num_blocks))
        width = conf.row_dim * (2 if self.update_fn.device_fn == DEV_ADAGRAD else 1)
        self.row_dim = conf.row_dim
        self.row_stride = _round_up(width, 32)
        if self.update_fn.device_fn == DEV_ADAGRAD:
            self.row_stride = 2 * _round_up(conf.row_dim, 32)   # kernel expects h at stride/2
        self.mode = PART_ORDERED if conf.is_ordered else PART_HASHED
        self.block_bytes = self.rows_per_block * self.row_stride * 4
        self.owners: List[Any] = list(block_owners) if block_owners is not None else \
            [b % self.world for b in range(self.num_blocks)]
        # local slabs
        self._local_ptr: Dict[int, int] = {}
        self._loc: Dict[int, Tuple[Any, int, int]] = {}   # block -> (owner key, seg, offset) for ALL blocks
        for b in range(self.num_blocks):
            if self.owners[b] == self.self_key:
                self._alloc_block(b)
        # pointer table lives in the arena too (so that a migrating peer can flip it remotely)
        self._ptr_tbl_ptr, tsi, toff = arena.alloc(self.num_blocks * 8)
        self.block_ptrs = arena.tensor(self._ptr_tbl_ptr, [self.num_blocks], torch.int64)
        self.block_ptrs.zero_()
        self._ptr_tbl_loc = (self.self_key, tsi, toff)
        self._peer_tbl_ptrs: Dict[Any, int] = {}
        if collective:
            self.apply_infos(self.group.all_gather_object(self.local_info()))
            if init:
                self.init_local_blocks()
            self.group.barrier()
        elif init:
            self.init_local_blocks()

    def _alloc_block(self, b: int) -> int:
        p, si, off = self.arena.alloc(self.block_bytes)
        self._local_ptr[b] = p
        self._loc[b] = (self.self_key, si, off)
        return p

    # ---- handle exchange & pointer table ----
    def local_info(self) -> Dict[str, Any]:
        return {"key": self.self_key,
                "blocks": {b: loc for b, loc in self._loc.items() if loc[0] == self.self_key},
                "handles": {si: self.arena.segment_info(si) for si in range(self.arena.num_segments())},
                "ptr_tbl": self._ptr_tbl_loc}

    def apply_infos(self, infos: Sequence[Dict[str, Any]]) -> None:
        """Map every peer's segments and fill the pointer table for the blocks they own."""
        for info in infos:
            r = info["key"]
            for si, h in info["handles"].items():
                self.arena.map_peer(r, si, h)
            for b, (rr, si, off) in info["blocks"].items():
                b = int(b)
                self._loc[b] = (rr, si, off)
                self.block_ptrs[b] = self.arena.resolve(rr, si, off)
            if info.get("ptr_tbl") is not None:
                _, tsi, toff = info["ptr_tbl"]
                self._peer_tbl_ptrs[r] = self.arena.resolve(r, tsi, toff)
        torch.cuda.synchronize(self.device)

    def ptrs_complete(self) -> bool:
        return bool((self.block_ptrs != 0).all().item())

    def _key_base_stride(self, b: int) -> Tuple[int, int]:
        return (b * self.rows_per_block, 1) if self.mode == PART_ORDERED else (b, self.num_blocks)

    def init_local_blocks(self) -> None:
        spec = self.update_fn.init_spec()
        for b, p in self._local_ptr.items():
            kb, ks = self._key_base_stride(b)
            if self.torch_dtype == torch.int32 or spec["kind"] == 0:
                self.local_block(b).zero_()
            else:
                self._n.init_slab(p, self.rows_per_block, self.row_stride, self.row_dim, spec["kind"],
                                  spec["a"], spec["b"], spec["seed"], kb, ks)
                if self.update_fn.device_fn == DEV_ADAGRAD:
                    self.local_block(b)[:, self.row_stride // 2:].zero_()

    # ---- views ----
    def view(self) -> Tuple[torch.Tensor, int, int, int, int]:
        """(block_ptrs, rows_per_block, row_stride, row_dim, mode): the kernel-side TableView."""
        return self.block_ptrs, self.rows_per_block, self.row_stride, self.row_dim, self.mode

    def local_block_ids(self) -> List[int]:
        return sorted(self._local_ptr)

    def local_block(self, b: int) -> torch.Tensor:
        return self.arena.tensor(self._local_ptr[b], [self.rows_per_block, self.row_stride], self.torch_dtype)

    def keys_of_block(self, b: int) -> np.ndarray:
        if self.mode == PART_ORDERED:
            lo = b * self.rows_per_block
            return np.arange(lo, min(self.conf.capacity, lo + self.rows_per_block), dtype=np.int64)
        return np.arange(b, self.conf.capacity, self.num_blocks, dtype=np.int64)

    def nbytes_local(self) -> int:
        return len(self._local_ptr) * self.block_bytes

    # ---- data plane ----
    def _check_keys(self, keys: torch.Tensor) -> None:
        """HB_CHECK_KEYS=1: validate keys before a kernel turns them into raw HBM addresses (costs a device sync per
        call; the loaders already reject out-of-range indices, so this is a debugging aid for library users)."""
        if _CHECK_KEYS and keys.numel():
            lo, hi = int(keys.min()), int(keys.max())
            if lo < 0 or hi >= self.conf.capacity:
                raise IndexError(f"table {self.conf.table_id}: key range [{lo}, {hi}] outside [0, {self.conf.capacity})")

    def pull(self, keys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K1: gather rows by key from whichever GPU owns them."""
        from ..ops import count_launch
        self._check_keys(keys)
        if out is None:
            out = torch.empty((keys.numel(), self.row_dim), dtype=self.torch_dtype, device=self.device)
        self._n.table_pull(*self.view(), keys, out)
        count_launch()
        return out

    def push(self, keys: torch.Tensor, deltas: torch.Tensor) -> None:
        """K2 fused with the table's update function, executed in the owner's HBM."""
        from ..ops import count_launch
        fn = self.update_fn.device_fn
        assert fn is not None, f"update function {self.update_fn.name} has no device implementation"
        p = self.update_fn.device_params()
        self._check_keys(keys)
        self._n.table_push(*self.view(), keys, deltas, fn, p["step"], p["lo"], p["hi"], p["eps"])
        count_launch()

    def push_sparse_i32(self, keys: torch.Tensor, topics: torch.Tensor, deltas: torch.Tensor) -> None:
        """K2+K5: LDA sparse (row, topic, delta) triples with clamp>=0 and nnz bookkeeping."""
        from ..ops import count_launch
        self._check_keys(keys)
        self._n.table_push_sparse_i32(*self.view(), keys, topics, deltas, self.row_dim - 1)
        count_launch()

    def put_rows(self, keys: torch.Tensor, values: torch.Tensor) -> None:
        from ..ops import count_launch
        self._check_keys(keys)
        self._n.table_push(*self.view(), keys, values.contiguous(), 3, 0.0, 0.0, 0.0, 0.0)
        count_launch()

    def scale_local(self, factor: float) -> None:
        """Owner-side decay of the local shard (L2 term of MLR applied at the owner)."""
        for b in self._local_ptr:
            self.local_block(b).mul_(factor)

    # ---- elasticity: block migration (collective over the group) ----
    def migrate_block(self, block: int, dst_rank: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Move `block` to `dst_rank`: slab copy by the receiver on a side stream, then the
        pointer flip on every GPU.  Callers quiesce accessors of THIS table at a mini-batch
        boundary (other jobs keep running; the copy overlaps their compute)."""
        from ..ops import count_launch
        src_rank = self.owners[block]
        if src_rank == dst_rank:
            return
        new_loc = None
        if self.rank == dst_rank:
            p, si, off = self.arena.alloc(self.block_bytes)
            new_loc = (self.self_key, si, off, self.arena.segment_info(si))
        infos = self.group.all_gather_object(new_loc)
        rr, si, off, handle = infos[dst_rank]
        if self.rank == dst_rank:
            src_va = int(self.block_ptrs[block].item())
            s = stream or torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._n.migrate_copy(src_va, p, self.block_bytes)
                count_launch()
            s.synchronize()
            self._local_ptr[block] = p
        self.group.barrier()          # data has landed at the receiver
        self.arena.map_peer(rr, si, handle)
        new_va = self.arena.resolve(rr, si, off)
        self.block_ptrs[block] = new_va   # ownership flip (local table of this GPU)
        self._loc[block] = (rr, si, off)
        self.owners[block] = dst_rank
        torch.cuda.synchronize(self.device)
        self.group.barrier()          # nobody references the old slab any more
        if self.rank == src_rank:
            self.arena.free(self._local_ptr.pop(block))

    # ---- non-collective migration primitives (driven by the job-server protocol) ----
    def recv_block(self, block: int, stream: Optional[torch.cuda.Stream] = None) -> Tuple[Any, int, int, Dict[str, Any]]:
        """Receiver half: allocate a slab and bulk-copy the block from its current owner's HBM
        (peer mapping) on a side stream.  Returns the new location to broadcast."""
        from ..ops import count_launch
        src_va = int(self.block_ptrs[block].item())
        p, si, off = self.arena.alloc(self.block_bytes)
        s = stream or torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._n.migrate_copy(src_va, p, self.block_bytes)
            count_launch()
        s.synchronize()
        self._pending_recv = getattr(self, "_pending_recv", {})
        self._pending_recv[block] = p
        return (self.self_key, si, off, self.arena.segment_info(si))

    def flip_block(self, block: int, new_owner: Any, loc: Tuple[Any, int, int, Any]) -> None:
        """Ownership flip on THIS GPU: map the new slab, rewrite the pointer-table entry, release
        the old slab when this GPU was the previous owner."""
        rr, si, off, handle = loc
        self.arena.map_peer(rr, si, handle)
        old_owner = self.owners[block]
        self.block_ptrs[block] = self.arena.resolve(rr, si, off)
        torch.cuda.synchronize(self.device)
        self._loc[block] = (rr, si, off)
        self.owners[block] = new_owner
        if new_owner == self.self_key:
            self._local_ptr[block] = getattr(self, "_pending_recv", {}).pop(block)
        elif old_owner == self.self_key and block in self._local_ptr:
            self.arena.free(self._local_ptr.pop(block))

    # ---- host import / export (checkpoint, tests) ----
    def export_block(self, b: int) -> Tuple[np.ndarray, np.ndarray]:
        keys = self.keys_of_block(b)
        rows = self.local_block(b)[: len(keys), : self.row_dim].cpu().numpy()
        return keys, rows

    def import_block(self, b: int, keys: np.ndarray, rows: np.ndarray) -> None:
        if len(keys) == 0:
            return
        keys = np.asarray(keys, dtype=np.int64)
        slots = (keys % self.rows_per_block) if self.mode == PART_ORDERED else (keys // self.num_blocks)
        blk = self.local_block(b)
        blk[torch.from_numpy(slots).to(self.device), : rows.shape[1]] = torch.from_numpy(np.ascontiguousarray(rows)).to(self.device)

    def to_dense(self) -> torch.Tensor:
        """Gather the whole table [capacity, row_dim] (tests / evaluation of small models)."""
        keys = torch.arange(self.conf.capacity, dtype=torch.int64, device=self.device)
        return self.pull(keys)

    def close(self) -> None:
        for p in self._local_ptr.values():
            self.arena.free(p)
        self._local_ptr.clear()
        self.arena.free(self._ptr_tbl_ptr)
