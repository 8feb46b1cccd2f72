"""Tracing: NVTX ranges on the GPU timeline + an optional chrome-trace JSON sink.

The reference wires HTrace span receivers (utils/trace/HTrace.java:30-57) that nothing calls
(SURVEY.md 5.1).  The B200 equivalent is NVTX ranges around PULL / COMP / PUSH (visible in ncu
/ nsys captures) and a process-local chrome-trace recorder for host-side spans.
"""
from __future__ import annotations

import contextlib
import json
import os
import threading
import time
from typing import Any, Dict, List, Optional

_events: List[Dict[str, Any]] = []
_lock = threading.Lock()
_enabled = bool(os.environ.get("HB_TRACE"))


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = flag


def _nvtx_push(name: str) -> bool:
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(name)
            return True
    except Exception:  # noqa: BLE001
        pass
    return False


def _nvtx_pop() -> None:
    try:
        import torch
        torch.cuda.nvtx.range_pop()
    except Exception:  # noqa: BLE001
        pass


@contextlib.contextmanager
def span(name: str, **args: Any):
    """`with span("PULL", job=...)`: NVTX range + chrome-trace complete event."""
    pushed = _nvtx_push(name) if _enabled else False
    t0 = time.perf_counter_ns()
    try:
        yield
    finally:
        if pushed:
            _nvtx_pop()
        if _enabled:
            t1 = time.perf_counter_ns()
            with _lock:
                _events.append({"name": name, "ph": "X", "ts": t0 / 1000.0, "dur": (t1 - t0) / 1000.0,
                                "pid": os.getpid(), "tid": threading.get_ident() % 100000, "args": args})


def dump(path: str) -> int:
    with _lock:
        evs = list(_events)
    with open(path, "w") as f:
        json.dump({"traceEvents": evs}, f)
    return len(evs)


def configure(conf) -> Optional[str]:
    """Apply the `-htrace_receiver_type/-host/-port` + `-base_dir` flags (reference: utils/trace/HTraceParameters).
    Returns the receiver type when tracing got enabled.  The spans are delivered by `flush_to_receiver()` (called
    by the launchers / the job server at exit)."""
    global _receiver
    kind = str(conf.get("htrace_receiver_type", "") or "").upper()
    if not kind:
        return None
    if kind not in ("STDOUT", "CHROME", "ZIPKIN"):
        raise ValueError(f"unknown -htrace_receiver_type {kind!r} (STDOUT | CHROME | ZIPKIN)")
    _receiver = {"kind": kind, "host": conf.get("htrace_receiver_host", "localhost"),
                 "port": int(conf.get("htrace_receiver_port", 9410)),
                 "dir": conf.get("base_dir", "/tmp/harmony_b200")}
    enable(True)
    return kind


_receiver: Optional[Dict[str, Any]] = None


def flush_to_receiver() -> Optional[str]:
    """Deliver the recorded spans to the configured receiver; returns where they went."""
    r = _receiver
    if r is None:
        return None
    with _lock:
        evs = list(_events)
        _events.clear()
    if r["kind"] == "STDOUT":
        for e in evs:
            print(json.dumps(e))
        return "stdout"
    if r["kind"] == "ZIPKIN":
        try:
            import urllib.request
            spans = [{"name": e["name"], "timestamp": int(e["ts"]), "duration": int(e["dur"]), "tags": e["args"],
                      "localEndpoint": {"serviceName": f"harmony_b200-{e['pid']}"}} for e in evs]
            req = urllib.request.Request(f"http://{r['host']}:{r['port']}/api/v2/spans", data=json.dumps(spans).encode(),
                                         headers={"Content-Type": "application/json"})
            urllib.request.urlopen(req, timeout=2).read()
            return f"zipkin://{r['host']}:{r['port']}"
        except Exception:  # noqa: BLE001  (collector unreachable: keep the spans on disk)
            pass
    os.makedirs(r["dir"], exist_ok=True)
    path = os.path.join(r["dir"], f"trace-{os.getpid()}.json")
    with open(path, "w") as f:
        json.dump({"traceEvents": evs}, f)
    return path


def clear() -> None:
    with _lock:
        _events.clear()


class Tracer:
    """Accumulating wall-clock timer (reference: dolphin/metric/Tracer.java:46-92)."""

    def __init__(self) -> None:
        self._begin: Optional[float] = None
        self._sum = 0.0
        self._count = 0

    def start_timer(self) -> None:
        self._begin = time.perf_counter()

    def record_time(self, n_elems: int = 1) -> None:
        if self._begin is not None:
            self._sum += time.perf_counter() - self._begin
            self._count += n_elems
            self._begin = None

    def total_elapsed_time(self) -> float:
        return self._sum

    def avg_time_per_elem(self) -> float:
        return self._sum / self._count if self._count else 0.0

    def reset_trace(self) -> None:
        self._begin, self._sum, self._count = None, 0.0, 0
