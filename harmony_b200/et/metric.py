"""Executor-side metric collector.

Reference: services/et/.../metric/MetricCollector.java:66-137 (periodic or manual flush of
tableToNumBlocks, bytesReceivedGetResp, countSentGetReq, hostname and encoded custom metrics),
MetricManager.java:46-56 (driver start/stop), driver/api/MetricReceiver.java:26-34.
"""
from __future__ import annotations

import threading
from typing import Any, List, Optional

from ..utils.misc import resolve_hostname


class MetricCollector:
    """Executor metric collector: table -> #blocks, GET-response bytes / counts, custom metrics; periodic or manual
    flush to the driver.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/metric/MetricCollector.java:66-137."""

    def __init__(self, executor) -> None:
        self._ex = executor
        self._custom: List[Any] = []
        self._lock = threading.Lock()
        self._timer: Optional[threading.Timer] = None
        self._period_s: float = 0.0
        self._running = False

    def start(self, flush_period_ms: int = 0) -> None:
        self._running = True
        self._period_s = flush_period_ms / 1000.0
        if self._period_s > 0:
            self._schedule()

    def _schedule(self) -> None:
        if not self._running:
            return
        self._timer = threading.Timer(self._period_s, self._tick)
        self._timer.daemon = True
        self._timer.start()

    def _tick(self) -> None:
        try:
            self.flush()
        finally:
            self._schedule()

    def stop(self) -> None:
        self._running = False
        if self._timer is not None:
            self._timer.cancel()
        self.flush()

    def add_custom_metric(self, metric: Any) -> None:
        with self._lock:
            self._custom.append(metric)

    def flush(self) -> None:
        with self._lock:
            custom, self._custom = self._custom, []
        count, nbytes = self._ex.sender.stat.get_and_reset()
        self._ex.send_to_driver({"type": "MetricReport", "executor_id": self._ex.executor_id,
                                 "table_to_num_blocks": self._ex.tables.get_table_to_num_blocks(),
                                 "count_sent_get_req": count, "bytes_received_get_resp": nbytes,
                                 "hostname": resolve_hostname(), "custom_metrics": custom})
