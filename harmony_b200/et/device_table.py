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
        # LDA count rows carry a hidden 32-word topic BITMAP in the last 32 words of the (padded) row: the sparse pull
        # reads it first and then only the 16 B chunks with non-zero counts (ops/csrc/table_ops.cu)
        self.bitmap_words = 32 if (self.update_fn.name == "lda_sparse" and conf.dtype != "float32"
                                   and 1 < conf.row_dim <= 1025) else 0
        self.row_stride = _round_up(width + self.bitmap_words, 32)
        self.bitmap_off = self.row_stride - 32 if self.bitmap_words else 0
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
        # + 2 guard words written by hb::row_ptr when a kernel meets an out-of-range key (count, last key)
        self._ptr_tbl_ptr, tsi, toff = arena.alloc((self.num_blocks + 2) * 8)
        self._view_ptrs = arena.tensor(self._ptr_tbl_ptr, [self.num_blocks + 2], torch.int64)
        self._view_ptrs.zero_()
        self.block_ptrs = self._view_ptrs[: self.num_blocks]
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
        return bool((self.block_ptrs != 0).all().item())   # (guard words excluded: block_ptrs is the [:num_blocks] view)

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
        return self._view_ptrs, self.rows_per_block, self.row_stride, self.row_dim, self.mode

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
    def check(self) -> None:
        """Raise if any kernel of this GPU met an out-of-range key since the last check.  The kernels bounds-check
        every key against the slab range in `hb::row_ptr` (count + clamp to key 0, no sync on the hot path); this
        reads the two guard words behind the pointer table (one small D2H copy) and resets them."""
        g = self._view_ptrs[self.num_blocks:].cpu()
        if int(g[0]) != 0:
            self._view_ptrs[self.num_blocks:].zero_()
            raise IndexError(f"table {self.conf.table_id}: {int(g[0])} out-of-range key(s) reached a kernel "
                             f"(last: {int(g[1])}, valid: [0, {self.conf.capacity})); they were clamped to key 0")


    def pull(self, keys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K1: gather rows by key from whichever GPU owns them."""
        from ..ops import count_launch
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
        self._n.table_push(*self.view(), keys, deltas, fn, p["step"], p["lo"], p["hi"], p["eps"])
        count_launch()

    def push_sparse_i32(self, keys: torch.Tensor, topics: torch.Tensor, deltas: torch.Tensor) -> None:
        """K2+K5: LDA sparse (row, topic, delta) triples with clamp>=0 and nnz bookkeeping."""
        from ..ops import count_launch
        self._n.table_push_sparse_i32(*self.view(), keys, topics, deltas, self.row_dim - 1, self.bitmap_off)
        count_launch()

    def pull_lda_sparse(self, keys: torch.Tensor, out: torch.Tensor, chunk_count: Optional[torch.Tensor] = None) -> None:
        """Sparse PULL of count rows through their topic bitmaps into the dense staging rows `out` [n, >= K + 1]."""
        from ..ops import count_launch
        assert self.bitmap_off > 0
        self._n.table_pull_lda_sparse(*self.view(), keys, out, self.row_dim - 1, self.bitmap_off, chunk_count)
        count_launch()

    def rebuild_bitmaps(self) -> None:
        """Owner-side: recompute the bitmaps of the local rows from their counts (drops stale bits)."""
        from ..ops import count_launch
        if not self.bitmap_off or not self._local_ptr:
            return
        ptrs = torch.tensor([self._local_ptr[b] for b in sorted(self._local_ptr)], dtype=torch.int64, device=self.device)
        self._n.lda_rebuild_bitmaps(ptrs, self.rows_per_block, self.row_stride, self.row_dim - 1, self.bitmap_off)
        count_launch()

    def put_rows(self, keys: torch.Tensor, values: torch.Tensor) -> None:
        """ASSIGN rows.  int32 tables (LDA counts) take a bit-exact integer path: converting the counts to float
        and storing them through the float kernel would write IEEE-754 bit patterns into the int32 slab."""
        from ..ops import count_launch
        if self.torch_dtype == torch.int32:
            vals = (values.round() if values.is_floating_point() else values).to(torch.int32)
            self._n.table_put_i32(*self.view(), keys, vals.contiguous(), self.bitmap_off)
        else:
            self._n.table_push(*self.view(), keys, values.float().contiguous(), 3, 0.0, 0.0, 0.0, 0.0)
        count_launch()

    def scale_local(self, factor: float) -> None:
        """Owner-side decay of the local shard (dense L2 term of MLR applied at the owner): ONE launch over the
        list of local slabs.  BSP trainers avoid even that with the lazy scale (models/mlr.py)."""
        from ..ops import count_launch
        if not self._local_ptr:
            return
        ptrs = torch.tensor([self._local_ptr[b] for b in sorted(self._local_ptr)], dtype=torch.int64,
                            device=self.device)
        self._n.scale_blocks(ptrs, self.block_bytes, float(factor))
        count_launch()

    # ---- elasticity: LIVE block migration (no accessor is stopped) -------------------------------------------
    # alloc (receiver) -> send (sender: ONE bulk-TMA launch for the whole list: old -> new over NVLink + old -> local
    # snapshot) -> flip (every GPU: one launch rewrites its pointer-table entries, then a fence = the kernels that
    # could still hold the old addresses have drained) -> drain (sender: new += old - snapshot, frees both).
    # Pushes that race with the copy land either in the new slab (after the flip) or in the old one; the latter are
    # forwarded by the drain, so additive update functions lose nothing (MigrationExecutor.java:134-295 gets the same
    # guarantee by latching the block; redirected stale accesses = RemoteAccessOpHandler.java:225-231).
    def alloc_recv(self, blocks: Sequence[int]) -> Dict[int, Tuple[Any, int, int, Dict[str, Any]]]:
        """Receiver: allocate the slabs of `blocks`; returns block -> (owner key, segment, offset, segment info)."""
        self._pending_recv = getattr(self, "_pending_recv", {})
        locs, done = {}, []
        try:
            for b in blocks:
                p, si, off = self.arena.alloc(self.block_bytes)
                self._pending_recv[int(b)] = p
                done.append(int(b))
                locs[int(b)] = (self.self_key, si, off, self.arena.segment_info(si))
        except Exception:
            self.abort_recv(done)
            raise
        return locs

    def abort_recv(self, blocks: Sequence[int]) -> None:
        for b in blocks:
            p = getattr(self, "_pending_recv", {}).pop(int(b), None)
            if p is not None:
                self.arena.free(p)

    def _resolve_locs(self, locs) -> Dict[int, int]:
        out = {}
        for b, (rr, si, off, handle) in locs.items():
            self.arena.map_peer(rr, si, handle)
            out[int(b)] = self.arena.resolve(rr, si, off)
        return out

    def _i64(self, xs) -> torch.Tensor:
        return torch.tensor(list(xs), dtype=torch.int64, device=self.device)

    def send_blocks(self, locs, stream: Optional[torch.cuda.Stream] = None) -> float:
        """Sender: bulk-copy the listed local blocks into their new slabs (peer HBM) and into local snapshots, in
        one launch on the side stream.  Returns the device time of the copy in milliseconds."""
        from ..ops import count_launch
        new_va = self._resolve_locs(locs)
        blocks = sorted(new_va)
        self._mig = getattr(self, "_mig", {})
        snaps = {b: self.arena.alloc(self.block_bytes)[0] for b in blocks}
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        s = stream or torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        try:
            with torch.cuda.stream(s):
                e0.record()
                self._n.migrate_bulk(self._i64(self._local_ptr[b] for b in blocks), self._i64(new_va[b] for b in blocks),
                                     self._i64(snaps[b] for b in blocks), self.block_bytes, err)
                e1.record()
                count_launch()
            s.synchronize()
            if int(err.item()) != 0:
                raise RuntimeError("bulk migration copy timed out on an mbarrier")
        except Exception:
            for p in snaps.values():             # nothing was flipped: the old slabs stay authoritative
                self.arena.free(p)
            raise
        for b in blocks:
            self._mig[b] = (self._local_ptr[b], snaps[b], new_va[b])
        return e0.elapsed_time(e1)

    def flip_blocks(self, locs, new_owner: Any, fence=None) -> None:
        """Every GPU: rewrite the pointer-table entries of the moved blocks (one launch, st.release), then fence:
        `fence()` must return once every kernel queued BEFORE the call has finished (default: device synchronize;
        the job server passes a per-stream event fence that does not wait for work queued later)."""
        from ..ops import count_launch
        new_va = self._resolve_locs(locs)
        blocks = sorted(new_va)
        cur = torch.cuda.current_stream(self.device)
        self._n.flip_blocks(self._view_ptrs, torch.tensor(blocks, dtype=torch.int32, device=self.device),
                            self._i64(new_va[b] for b in blocks))
        count_launch()
        cur.synchronize()                        # the new addresses are in HBM: kernels starting from now on use them
        if fence is not None:
            fence()                              # every kernel that may have read an old address has finished
        else:
            torch.cuda.synchronize(self.device)
        self.flip_count = getattr(self, "flip_count", 0) + 1
        for b in blocks:
            rr, si, off, _h = locs[b] if b in locs else locs[str(b)]
            self._loc[b] = (rr, si, off)
            self.owners[b] = new_owner
            if new_owner == self.self_key:
                self._local_ptr[b] = self._pending_recv.pop(b)
            elif b in self._local_ptr:
                self._local_ptr.pop(b)          # the old slab now only lives in self._mig until the drain

    def drain_blocks(self, blocks: Sequence[int], stream: Optional[torch.cuda.Stream] = None) -> int:
        """Sender, after EVERY GPU has flipped: forward what still landed in the old slabs (new += old - snapshot),
        then free the old slabs and the snapshots.  Returns the number of forwarded elements."""
        from ..ops import count_launch
        blocks = [int(b) for b in blocks if int(b) in getattr(self, "_mig", {})]
        if not blocks:
            return 0
        fw = torch.zeros(1, dtype=torch.int64, device=self.device)
        s = stream or torch.cuda.current_stream(self.device)
        with torch.cuda.stream(s):
            from .update_function import DEV_SGD_CLAMP
            p = self.update_fn.device_params()
            clamp = self.update_fn.device_fn == DEV_SGD_CLAMP       # bounded rows stay inside [lo, hi] after the forward
            self._n.migrate_drain(self._i64(self._mig[b][0] for b in blocks), self._i64(self._mig[b][1] for b in blocks),
                                  self._i64(self._mig[b][2] for b in blocks), self.block_bytes,
                                  self.torch_dtype == torch.int32, fw, clamp, p["lo"], p["hi"])
            count_launch()
        s.synchronize()
        for b in blocks:
            old, snap, _new = self._mig.pop(b)
            self.arena.free(old)
            self.arena.free(snap)
        return int(fw.item())

    def migrate_blocks(self, blocks: Sequence[int], dst_rank: Any, stream: Optional[torch.cuda.Stream] = None
                       ) -> Dict[str, float]:
        """SPMD form of the live move (collective over the group; ONE object exchange for the whole list)."""
        blocks = [int(b) for b in blocks if self.owners[int(b)] != dst_rank]
        if not blocks:
            return {"blocks": 0}
        src_rank = self.owners[blocks[0]]
        assert all(self.owners[b] == src_rank for b in blocks), "one sender per migrate_blocks call"
        locs = self.alloc_recv(blocks) if self.self_key == dst_rank else None
        locs = next(x for x in self.group.all_gather_object(locs) if x is not None)
        copy_ms = self.send_blocks(locs, stream) if self.self_key == src_rank else 0.0
        self.group.barrier()                 # the copies have landed
        self.flip_blocks(locs, dst_rank)
        self.group.barrier()                 # every GPU flipped and fenced
        forwarded = self.drain_blocks(blocks, stream) if self.self_key == src_rank else 0
        self.group.barrier()
        return {"blocks": len(blocks), "bytes": len(blocks) * self.block_bytes, "copy_ms": copy_ms,
                "forwarded_elements": forwarded}

    def migrate_block(self, block: int, dst_rank: Any, stream: Optional[torch.cuda.Stream] = None) -> None:
        self.migrate_blocks([block], dst_rank, stream)

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
