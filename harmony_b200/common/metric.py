"""Generic metric tracker framework.

Reference: common/.../metric/MetricsCollector.java:35-110 (start/stop aggregates the registered
MetricTrackers and hands the result to a MetricsHandler), TimeMetricTracker, MemoryMetricTracker
(periodic sampling), GCMetricTracker, InsertableMetricTracker, common/src/main/avro/metric.avsc.
On a GPU box "memory" means HBM (torch.cuda.memory_stats) next to host RSS, and GC is replaced by
CUDA-event timers.
"""
from __future__ import annotations

import threading
import time
from typing import Callable, Dict, List, Optional

from ..utils.misc import device_memory_mb, used_memory_mb


class MetricTracker:
    """start() / stop() -> metric dict SPI of the generic metric framework.  Reference:
    common/src/main/java/edu/snu/cay/common/metric/MetricTracker.java."""

    def start(self) -> None:
        pass

    def stop(self) -> Dict[str, float]:
        return {}

    def close(self) -> None:
        pass


class TimeMetricTracker(MetricTracker):
    """Wall-clock time between start and stop.  Reference:
    common/src/main/java/edu/snu/cay/common/metric/TimeMetricTracker.java."""

    KEY = "METRIC_WALL_TIME_SEC"

    def start(self) -> None:
        self._t0 = time.perf_counter()

    def stop(self) -> Dict[str, float]:
        return {self.KEY: time.perf_counter() - self._t0}


class MemoryMetricTracker(MetricTracker):
    """Samples host RSS and device memory every `period_s`; reports max and mean of the bracketed window.  One
    sampler thread serves every start / stop window of the tracker (an epoch of a fast job is shorter than the
    thread start + join it used to cost); it retires by itself after a few idle periods."""

    IDLE_PERIODS = 4

    def __init__(self, period_s: float = 0.1) -> None:
        self.period_s = period_s
        self._samples: List[tuple] = []
        self._lock = threading.Lock()
        self._active = False
        self._closed = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._last: Optional[tuple] = None

    def _run(self) -> None:
        idle = 0
        while not self._closed.wait(self.period_s):
            with self._lock:
                if not self._active:
                    idle += 1
                    if idle >= self.IDLE_PERIODS:
                        self._thread = None
                        return
                    continue
                idle = 0
            sample = (used_memory_mb(), device_memory_mb())
            with self._lock:
                if self._active:
                    self._samples.append(sample)

    def start(self) -> None:
        with self._lock:
            self._samples = []
            self._active = True
            if self._thread is None and not self._closed.is_set():
                self._thread = threading.Thread(target=self._run, daemon=True, name="memory-metric")
                self._thread.start()

    def stop(self) -> Dict[str, float]:
        with self._lock:
            self._active = False
            s, self._samples = self._samples, []
        if not s:
            # window shorter than the sampling period: one sample now, reused by windows that close within a period
            now = time.perf_counter()
            if self._last is None or now - self._last[0] >= self.period_s:
                self._last = (now, (used_memory_mb(), device_memory_mb()))
            s = [self._last[1]]
        host, dev = [a for a, _ in s], [b for _, b in s]
        return {"METRIC_MEMORY_MAX_MB": max(host), "METRIC_MEMORY_AVG_MB": sum(host) / len(host),
                "METRIC_DEVICE_MEMORY_MAX_MB": max(dev), "METRIC_DEVICE_MEMORY_AVG_MB": sum(dev) / len(dev)}

    def close(self) -> None:
        """Retire the sampler thread now (it would also retire by itself)."""
        self._closed.set()


class DeviceTimeMetricTracker(MetricTracker):
    """cudaEvent-based device time of the bracketed region (replaces the JVM GC tracker)."""
    KEY = "METRIC_DEVICE_TIME_SEC"

    def start(self) -> None:
        import torch
        self._ok = torch.cuda.is_available()
        if self._ok:
            self._e0, self._e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self._e0.record()

    def stop(self) -> Dict[str, float]:
        if not self._ok:
            return {self.KEY: 0.0}
        import torch
        self._e1.record()
        torch.cuda.synchronize()
        return {self.KEY: self._e0.elapsed_time(self._e1) / 1e3}


class InsertableMetricTracker(MetricTracker):
    """User-inserted key/value metrics.  Reference:
    common/src/main/java/edu/snu/cay/common/metric/InsertableMetricTracker.java."""

    def __init__(self) -> None:
        self._vals: Dict[str, float] = {}

    def put(self, key: str, value: float) -> None:
        self._vals[key] = value

    def stop(self) -> Dict[str, float]:
        out, self._vals = self._vals, {}
        return out


class MetricsCollector:
    """Aggregates the registered trackers' metrics between start() and stop() and hands them to a handler.
    Reference: common/src/main/java/edu/snu/cay/common/metric/MetricsCollector.java."""

    def __init__(self, handler: Optional[Callable[[Dict[str, float]], None]] = None) -> None:
        self._trackers: List[MetricTracker] = []
        self._handler = handler
        self.last: Dict[str, float] = {}

    def register_trackers(self, trackers: List[MetricTracker]) -> None:
        self._trackers.extend(trackers)

    def start(self) -> None:
        for t in self._trackers:
            t.start()

    def stop(self) -> Dict[str, float]:
        out: Dict[str, float] = {}
        for t in self._trackers:
            out.update(t.stop())
        self.last = out
        if self._handler is not None:
            self._handler(out)
        return out

    def close(self) -> None:
        for t in self._trackers:
            t.close()
