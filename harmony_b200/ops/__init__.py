"""Native ops of harmony_b200 (hand-written sm_100a CUDA + peer arena).

`native()` returns the compiled extension and FAILS LOUDLY when it is missing on a GPU box:
there is no silent PyTorch fallback for the device data plane.
"""
from __future__ import annotations

from . import build as _build

_LAUNCHES = 0   # number of our kernels launched (bench.py reports it as gpu_launches)


def native():
    return _build.load()


def native_available() -> bool:
    return _build.available()


def native_or_none():
    """The extension if it is built and loadable, else None -- for the HOST-side helpers (text parsers, batch
    grouping) that have a NumPy twin.  The device data plane uses `native()` and never falls back."""
    if not _build.available():
        return None
    try:
        return _build.load()
    except (ImportError, OSError):
        return None


def count_launch(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def launches() -> int:
    return _LAUNCHES


def reset_launches() -> None:
    global _LAUNCHES
    _LAUNCHES = 0
