"""Peer memory plumbing: one process per GPU, every process maps every peer's table arena.

`PeerGroup` is the bootstrap group (rank / world / object exchange) -- `torch.distributed` when a
process group exists (NCCL on GPUs, gloo on CPU), a trivial single-member group otherwise.
`PeerArena` is a segmented HBM arena whose segments are mapped by all peers, so that any block slab
has a valid virtual address in every process: the kernels then read / reduce into the owner's HBM
over NVLink 5 / NVSwitch (no NCCL on the data path).  Two segment kinds:
  * "vmm" (default on a GPU that supports it): cuMemCreate physical memory, exported as a POSIX fd
    that travels over a unix socket (`FdServer`), mapped by the peers with cuMemMap.  Only VMM memory can
    be bound to a multicast object, so this is what makes the NVLS data plane possible;
  * "ipc": cudaMalloc + CUDA IPC handles (round-1 path, kept as the fallback).
`SymmetricRegion` is a same-offset buffer on every rank with all peers' unicast pointers AND, when the
fabric supports it, one multicast (NVLS) pointer: `multimem.st` on it broadcasts in the switch,
`multimem.ld_reduce` sums the replicas in the switch (ops/csrc/vmm.cpp, ops/csrc/mlr_ship_ops.cu).
"""
from __future__ import annotations

import os
import socket
import struct
import threading
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch


class FdServer:
    """Per-process unix-socket server that hands out POSIX fds of VMM allocations (SCM_RIGHTS).

    cuMemExportToShareableHandle yields a file descriptor, which only means something inside the exporting
    process; the importer connects to `\\0hb-fd-<pid>` (abstract namespace: no file, dies with the process),
    names the allocation and receives a duplicate of the fd.  One daemon thread, started on first use."""

    _inst: Optional["FdServer"] = None
    _lock = threading.Lock()

    def __init__(self) -> None:
        self._providers: Dict[str, Callable[[], int]] = {}
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._sock.bind(self.address(os.getpid()))
        self._sock.listen(64)
        threading.Thread(target=self._serve, name="hb-fd-server", daemon=True).start()

    @staticmethod
    def address(pid: int) -> str:
        return "\0hb-fd-%d" % pid

    @classmethod
    def get(cls) -> "FdServer":
        with cls._lock:
            if cls._inst is None:
                cls._inst = FdServer()
            return cls._inst

    def register(self, key: str, provider: Callable[[], int]) -> None:
        self._providers[key] = provider

    def unregister(self, key: str) -> None:
        self._providers.pop(key, None)

    def _serve(self) -> None:
        while True:
            try:
                conn, _ = self._sock.accept()
            except OSError:
                return
            try:
                n = struct.unpack("<I", conn.recv(4, socket.MSG_WAITALL))[0]
                key = conn.recv(n, socket.MSG_WAITALL).decode()
                prov = self._providers.get(key)
                if prov is None:
                    conn.sendall(b"\x00")
                else:
                    fd = int(prov())
                    try:
                        socket.send_fds(conn, [b"\x01"], [fd])
                    finally:
                        os.close(fd)
            except Exception:  # noqa: BLE001  (a bad client must not kill the server)
                pass
            finally:
                conn.close()

    @staticmethod
    def fetch(pid: int, key: str, timeout: float = 30.0) -> int:
        """Receive a duplicate of the fd registered as `key` in process `pid` (caller closes it)."""
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.settimeout(timeout)
        try:
            c.connect(FdServer.address(pid))
            kb = key.encode()
            c.sendall(struct.pack("<I", len(kb)) + kb)
            msg, fds, _flags, _addr = socket.recv_fds(c, 1, 1)
            if msg != b"\x01" or not fds:
                raise KeyError(f"process {pid} exports no allocation named {key!r}")
            return fds[0]
        finally:
            c.close()


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
    """Segmented device arena; each segment is a VMM allocation (fd-shareable, multicast-bindable) or, as the
    fallback, one cudaMalloc exported with cudaIpcGetMemHandle."""

    DEFAULT_SEGMENT = 1 << 30

    def __init__(self, device: int, group: Optional[PeerGroup] = None, segment_bytes: Optional[int] = None,
                 self_key: Any = None, kind: Optional[str] = None) -> None:
        from ..ops import native
        self._n = native()
        self.device = int(device)
        kind = kind or os.environ.get("HB_ARENA_KIND", "auto")
        if kind == "auto":
            kind = "vmm" if self._n.vmm_supported(self.device)[0] else "ipc"
        assert kind in ("vmm", "ipc"), kind
        self.kind = kind
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
        self._vmm_imports: set = set()

    # ---- local allocation ----
    def alloc(self, nbytes: int) -> Tuple[int, int, int]:
        """Returns (device_ptr, segment_index, offset_in_segment)."""
        for si, (aid, _sz) in enumerate(self._segments):
            p = self._n.arena_alloc(aid, nbytes)
            if p:
                self._alloc_seg[p] = si
                return p, si, p - aid
        seg_bytes = max(self.segment_bytes, nbytes)
        si = self.new_segment(seg_bytes)
        aid = self._segments[si][0]
        p = self._n.arena_alloc(aid, nbytes)
        if not p:
            raise MemoryError(f"arena segment of {seg_bytes} bytes cannot hold {nbytes} bytes")
        self._alloc_seg[p] = si
        return p, si, p - aid

    def new_segment(self, nbytes: int, align: int = 0) -> int:
        """Create one more segment (VMM: physical size rounded up to the allocation granularity and `align`)."""
        aid = self._n.arena_create(self.device, int(nbytes), 1 if self.kind == "vmm" else 0, int(align))
        cap = int(self._n.arena_stats(aid)[0])
        self._segments.append((aid, cap))
        if self.kind == "vmm":
            FdServer.get().register(self._fd_key(aid), lambda a=aid: self._n.vmm_export_fd(a))
        return len(self._segments) - 1

    @staticmethod
    def _fd_key(base: int) -> str:
        return "seg:%x" % base

    def alloc_in(self, seg_index: int, nbytes: int) -> Tuple[int, int, int]:
        """Allocate from one specific segment (symmetric regions own a whole segment)."""
        aid = self._segments[seg_index][0]
        p = self._n.arena_alloc(aid, nbytes)
        if not p:
            raise MemoryError(f"segment {seg_index} cannot hold {nbytes} bytes")
        self._alloc_seg[p] = seg_index
        return p, seg_index, p - aid

    def free(self, ptr: int) -> None:
        si = self._alloc_seg.pop(ptr)
        self._n.arena_free(self._segments[si][0], ptr)

    def segment_handle(self, seg_index: int) -> bytes:
        return bytes(self._n.ipc_export(self._segments[seg_index][0]))

    def segment_info(self, seg_index: int) -> Dict[str, Any]:
        """Everything a peer needs to map this segment: the fd-server key (VMM) or the IPC handle for other
        processes, the raw base pointer for executors living in the SAME process."""
        base, nbytes = self._segments[seg_index]
        info = {"kind": self.kind, "pid": os.getpid(), "device": self.device, "base": base, "bytes": nbytes}
        if self.kind == "vmm":
            info["fd_key"] = self._fd_key(base)
        else:
            info["handle"] = self.segment_handle(seg_index)
        return info

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
                if handle.get("kind") == "vmm":     # process-wide VA: grant this GPU access to the peer's memory
                    self._n.vmm_grant(handle["base"], self.device)
                else:
                    self._n.enable_peer_access(self.device, handle["device"])
            base = handle["base"]
            self._same_process.add(key)
        elif isinstance(handle, dict) and handle.get("kind") == "vmm":
            fd = FdServer.fetch(handle["pid"], handle["fd_key"])
            try:
                base = self._n.vmm_import_fd(fd, handle["bytes"], self.device)
            finally:
                os.close(fd)
            self._vmm_imports.add(base)
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
                    if base in self._vmm_imports:
                        self._n.vmm_release(base)
                    else:
                        self._n.ipc_close(base)
                except Exception:  # noqa: BLE001
                    pass
        self._peer_base.clear()
        self._vmm_imports.clear()
        for aid, _ in self._segments:
            if self.kind == "vmm":
                FdServer.get().unregister(self._fd_key(aid))
            self._n.arena_destroy(aid)
        self._segments.clear()


class SymmetricRegion:
    """`nbytes` at the SAME offset of a dedicated segment on every rank (collective constructor).

    `local_ptr`  this rank's buffer, `peer_ptrs[r]` rank r's buffer as mapped here (unicast over NVLink),
    `mc_ptr`     the multicast (NVLS) address of the buffer or 0 when the arena is not VMM / the fabric has no
                 multicast / any rank failed to join (the decision is collective, nobody is left waiting in a bind).
    """

    def __init__(self, arena: PeerArena, group: Optional[PeerGroup], nbytes: int, multicast: bool = True) -> None:
        self.arena, self.group = arena, group or arena.group
        self._n = arena._n
        rank, world = self.group.rank, self.group.world
        self.rank, self.world = rank, world
        want_mc = bool(multicast and world > 1 and arena.kind == "vmm" and
                       os.environ.get("HB_NVLS", "1") != "0" and self._n.vmm_supported(arena.device)[1])
        flags = self.group.all_gather_object(want_mc)
        want_mc = all(flags)
        gran = 2 << 20
        if want_mc:
            try:
                gran = int(self._n.mc_granularity(world, 1 << 30, True))
            except Exception:  # noqa: BLE001
                want_mc = False
        self.nbytes = (int(nbytes) + gran - 1) // gran * gran
        self.seg = arena.new_segment(self.nbytes, gran)
        self.local_ptr = arena.segment_base(self.seg)
        # the whole segment is the region: take it from the segment's allocator so nobody else lands in it
        p, si, off = arena.alloc_in(self.seg, self.nbytes)
        assert p == self.local_ptr and off == 0
        infos = self.group.all_gather_object((rank, self.seg, arena.segment_info(self.seg)))
        self.peer_ptrs = [0] * world
        for r, si, h in infos:
            arena.map_peer(r, si, h)
            self.peer_ptrs[r] = arena.resolve(r, si, 0)
        self.mc_ptr = 0
        self._mc_id = None
        if want_mc:
            try:
                self._setup_multicast(infos)
            finally:
                if getattr(self, "_mc_fd", None) is not None:      # rank 0: the exported fd has served its purpose
                    FdServer.get().unregister("mc:%x" % self.local_ptr)
                    try:
                        self._n.close_fd(self._mc_fd)
                    except Exception:  # noqa: BLE001
                        pass
                    self._mc_fd = None
        self.group.barrier()

    def _stage_ok(self, ok: bool) -> bool:
        return all(self.group.all_gather_object(bool(ok)))

    def _setup_multicast(self, infos) -> None:
        """cuMulticastCreate on rank 0 -> fd to everyone -> add device -> bind own segment -> map.  Every stage is
        followed by a collective vote so that a failure anywhere downgrades ALL ranks to the unicast path."""
        n, dev, rank, world = self._n, self.arena.device, self.rank, self.world
        mc_id, ok = None, True
        key = "mc:%x" % self.local_ptr
        try:
            if rank == 0:
                mc_id, fd = n.mc_create(dev, world, self.nbytes)
                self._mc_fd = fd
                FdServer.get().register(key, lambda f=fd: os.dup(f))
        except Exception as e:  # noqa: BLE001
            ok, self.mc_error = False, repr(e)
        root = self.group.all_gather_object((os.getpid(), key, ok))[0]
        if not root[2]:
            return
        try:
            if rank != 0:
                fd = FdServer.fetch(root[0], root[1])
                try:
                    mc_id = n.mc_import(fd, dev, world, self.nbytes)
                finally:
                    os.close(fd)
            n.mc_add_device(mc_id, dev)
        except Exception as e:  # noqa: BLE001
            ok, self.mc_error = False, repr(e)
        if not self._stage_ok(ok):
            return
        try:
            n.mc_bind(mc_id, 0, self.local_ptr, 0, self.nbytes)
        except Exception as e:  # noqa: BLE001
            ok, self.mc_error = False, repr(e)
        if not self._stage_ok(ok):
            return
        try:
            va = n.mc_map(mc_id, dev)
        except Exception as e:  # noqa: BLE001
            ok, va, self.mc_error = False, 0, repr(e)
        if not self._stage_ok(ok):
            return
        self.mc_ptr, self._mc_id = int(va), mc_id

    def tensor(self, offset: int, shape, dtype=torch.float32) -> torch.Tensor:
        return self.arena.tensor(self.local_ptr + offset, shape, dtype)

    def close(self) -> None:
        if self._mc_id is not None:
            torch.cuda.synchronize(self.arena.device)
            self._n.mc_unmap(self.mc_ptr, self.nbytes)
            self._n.mc_destroy(self._mc_id, self.arena.device)
            self._mc_id, self.mc_ptr = None, 0
