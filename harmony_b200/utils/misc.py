"""Small helpers (reference: utils/{LongRangeUtils,Tuple3,MemoryUtils,HostnameResolver}.java)."""
from __future__ import annotations

import os
import sys
import socket
from dataclasses import dataclass
from typing import Generic, List, Sequence, Tuple, TypeVar

A = TypeVar("A")
B = TypeVar("B")
C = TypeVar("C")


@dataclass(frozen=True)
class Tuple3(Generic[A, B, C]):
    """Immutable triple.  Reference: utils/src/main/java/edu/snu/cay/utils/Tuple3.java."""

    first: A
    second: B
    third: C


def generate_dense_long_ranges(keys: Sequence[int]) -> List[Tuple[int, int]]:
    """Sorted keys -> minimal list of closed [lo, hi] ranges (LongRangeUtils.generateDenseLongRanges)."""
    out: List[Tuple[int, int]] = []
    for k in sorted(set(int(x) for x in keys)):
        if out and out[-1][1] + 1 == k:
            out[-1] = (out[-1][0], k)
        else:
            out.append((k, k))
    return out


def get_num_units(ranges: Sequence[Tuple[int, int]]) -> int:
    return sum(hi - lo + 1 for lo, hi in ranges)


_PROC = None


def used_memory_mb() -> float:
    """Resident set size of this process (the psutil handle is cached: building one reads /proc/<pid>/stat)."""
    global _PROC
    try:
        if _PROC is None or _PROC.pid != os.getpid():
            import psutil
            _PROC = psutil.Process(os.getpid())
        return _PROC.memory_info().rss / (1 << 20)
    except Exception:  # noqa: BLE001
        return 0.0


def device_memory_mb() -> float:
    """HBM allocated through torch in this process; 0 when torch was never loaded (a host-backend executor must not
    pay the torch import -- and the import lock -- for a memory sample)."""
    torch = sys.modules.get("torch")
    try:
        if torch is not None and torch.cuda.is_available():
            return torch.cuda.memory_allocated() / (1 << 20)
    except Exception:  # noqa: BLE001
        pass
    return 0.0


_HOSTNAME = None


def resolve_hostname() -> str:
    """Cached: the metric collector stamps every per-mini-batch report with it."""
    global _HOSTNAME
    if _HOSTNAME is None:
        try:
            _HOSTNAME = socket.gethostname()
        except Exception:  # noqa: BLE001
            _HOSTNAME = "localhost"
    return _HOSTNAME


def murmur3_32(data: bytes, seed: int = 0) -> int:
    """MurmurHash3 x86_32 (the hash used by the hashed block partitioner)."""
    c1, c2 = 0xcc9e2d51, 0x1b873593
    h = seed & 0xffffffff
    n = len(data)
    rounded = n & ~3
    for i in range(0, rounded, 4):
        k = int.from_bytes(data[i:i + 4], "little")
        k = (k * c1) & 0xffffffff
        k = ((k << 15) | (k >> 17)) & 0xffffffff
        k = (k * c2) & 0xffffffff
        h ^= k
        h = ((h << 13) | (h >> 19)) & 0xffffffff
        h = (h * 5 + 0xe6546b64) & 0xffffffff
    k = 0
    tail = n & 3
    if tail == 3:
        k ^= data[rounded + 2] << 16
    if tail >= 2:
        k ^= data[rounded + 1] << 8
    if tail >= 1:
        k ^= data[rounded]
        k = (k * c1) & 0xffffffff
        k = ((k << 15) | (k >> 17)) & 0xffffffff
        k = (k * c2) & 0xffffffff
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85ebca6b) & 0xffffffff
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & 0xffffffff
    h ^= h >> 16
    return h
