"""Peer memory plumbing: one process per GPU, every process maps every peer's table arena.

`PeerGroup` is the bootstrap group (rank / world / object exchange) -- `torch.distributed` when a
process group exists (NCCL on GPUs, gloo on CPU), a trivial single-member group otherwise.
`PeerArena` is a segmented HBM arena whose segments are exported through CUDA IPC and opened by
all peers, so that any block slab has a valid virtual address in every process: the kernels then
read / reduce into the owner's HBM over NVLink 5 / NVSwitch (no NCCL on the data path).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import torch


class PeerGroup:
    """Bootstrap group. NCCL/gloo are used ONLY to exchange handles and for baselines/barriers."""

    def __init__(self, rank: int = 0, world: int = 1, pg: Any = None) -> None:
        self.rank, self.world, self.pg = rank, world, pg

    @staticmethod
    def from_env() -> "PeerGroup":
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return PeerGroup(dist.get_rank(), dist.get_world_size(), dist.group.WORLD)
        return PeerGroup(0, 1, None)

    def all_gather_object(self, obj: Any) -> List[Any]:
        if self.world == 1 or self.pg is None:
            return [obj]
        import torch.distributed as dist
        out: List[Any] = [None] * self.world
        dist.all_gather_object(out, obj, group=self.pg)
        return out

    def barrier(self) -> None:
        if self.world > 1 and self.pg is not None:
            import torch.distributed as dist
            dist.barrier(group=self.pg)


class PeerArena:
    """Segmented device arena; each segment = one cudaMalloc exported with cudaIpcGetMemHandle."""

    DEFAULT_SEGMENT = 1 << 30

    def __init__(self, device: int, group: Optional[PeerGroup] = None, segment_bytes: Optional[int] = None,
                 self_key: Any = None) -> None:
        from ..ops import native
        self._n = native()
        self.device = int(device)
        self.group = group or PeerGroup()
        # identity of this arena among its peers: SPMD rank by default, executor id in the job server
        self.self_key = self.group.rank if self_key is None else self_key
        self.segment_bytes = int(segment_bytes or os.environ.get("HB_ARENA_SEGMENT_BYTES", self.DEFAULT_SEGMENT))
        self._segments: List[Tuple[int, int]] = []            # (arena_id/base_ptr, bytes)
        self._alloc_seg: Dict[int, int] = {}                  # ptr -> segment index
        # peer mappings: (rank, seg_index) -> base VA in this process
        self._peer_base: Dict[Tuple[int, int], int] = {}
        self._peer_handles: Dict[Tuple[int, int], bytes] = {}
        self._same_process: set = set()

    # ---- local allocation ----
    def alloc(self, nbytes: int) -> Tuple[int, int, int]:
        """Returns (device_ptr, segment_index, offset_in_segment)."""
        for si, (aid, _sz) in enumerate(self._segments):
            p = self._n.arena_alloc(aid, nbytes)
            if p:
                self._alloc_seg[p] = si
                return p, si, p - aid
        seg_bytes = max(self.segment_bytes, nbytes)
        aid = self._n.arena_create(self.device, seg_bytes)
        self._segments.append((aid, seg_bytes))
        si = len(self._segments) - 1
        p = self._n.arena_alloc(aid, nbytes)
        if not p:
            raise MemoryError(f"arena segment of {seg_bytes} bytes cannot hold {nbytes} bytes")
        self._alloc_seg[p] = si
        return p, si, p - aid

    def free(self, ptr: int) -> None:
        si = self._alloc_seg.pop(ptr)
        self._n.arena_free(self._segments[si][0], ptr)

    def segment_handle(self, seg_index: int) -> bytes:
        return bytes(self._n.ipc_export(self._segments[seg_index][0]))

    def segment_info(self, seg_index: int) -> Dict[str, Any]:
        """Everything a peer needs to map this segment: IPC handle for other processes, raw base
        pointer for executors living in the SAME process (CUDA IPC cannot re-open its own handle)."""
        return {"handle": self.segment_handle(seg_index), "pid": os.getpid(), "device": self.device,
                "base": self._segments[seg_index][0]}

    def segment_base(self, seg_index: int) -> int:
        return self._segments[seg_index][0]

    def num_segments(self) -> int:
        return len(self._segments)

    def tensor(self, ptr: int, shape, dtype=torch.float32, strides=None) -> torch.Tensor:
        code = {torch.float32: 0, torch.int32: 1, torch.int64: 2, torch.uint8: 3}[dtype]
        shape = list(shape)
        if strides is None:
            strides, acc = [], 1
            for s in reversed(shape):
                strides.append(acc)
                acc *= s
            strides = list(reversed(strides))
        return self._n.tensor_from_ptr(ptr, shape, list(strides), code, self.device)

    # ---- peer mapping ----
    def map_peer(self, rank: Any, seg_index: int, handle: Any) -> int:
        """Map (once) segment `seg_index` of peer `rank`; returns its base VA in this process.
        `handle` is raw IPC bytes or a `segment_info()` dict."""
        key = (rank, seg_index)
        if key in self._peer_base:
            return self._peer_base[key]
        if rank == self.self_key:
            base = self._segments[seg_index][0]
        elif isinstance(handle, dict) and handle.get("pid") == os.getpid():
            if handle["device"] != self.device:
                self._n.enable_peer_access(self.device, handle["device"])
            base = handle["base"]
            self._same_process.add(key)
        else:
            raw = handle["handle"] if isinstance(handle, dict) else handle
            base = self._n.ipc_open(raw, self.device)
        self._peer_base[key] = base
        self._peer_handles[key] = handle
        return base

    def resolve(self, rank: Any, seg_index: int, offset: int, handle: Any = None) -> int:
        if rank == self.self_key:
            return self._segments[seg_index][0] + offset
        key = (rank, seg_index)
        if key not in self._peer_base:
            if handle is None:
                raise KeyError(f"segment {seg_index} of rank {rank} is not mapped")
            self.map_peer(rank, seg_index, handle)
        return self._peer_base[key] + offset

    def close(self) -> None:
        for (rank, _si), base in self._peer_base.items():
            if rank != self.self_key and (rank, _si) not in self._same_process:
                try:
                    self._n.ipc_close(base)
                except Exception:  # noqa: BLE001
                    pass
        self._peer_base.clear()
        for aid, _ in self._segments:
            self._n.arena_destroy(aid)
        self._segments.clear()
