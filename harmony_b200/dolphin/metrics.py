"""Dolphin metric records + the master-side metric store.

Reference: jobserver/src/main/avro/metrics.avsc:25-244 (BatchMetrics{batchTimeSec,
batchCompTimeSec, batchPullTimeSec, batchPushTimeSec, dataProcessingRate, numBatchDataInstances,
...}, EpochMetrics, WorkerMetrics, ServerMetrics), dolphin/metric/ETDolphinMetricReceiver.java:64-201,
dolphin/metric/MetricManager.java:113-160,230-257,262-376 (validation: skip first N, skip the
first metric after a reconfiguration, block counts must match the driver's view; pause/resume).
"""
from __future__ import annotations

import threading
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional


@dataclass
class BatchMetrics:
    """Per-mini-batch worker metrics (times, processing rate, custom metrics).  Reference:
    jobserver/src/main/avro/metrics.avsc (BatchMetrics)."""

    job_id: str
    worker_id: str
    epoch_idx: int
    batch_idx: int
    num_batch_data_instances: int
    batch_time_sec: float
    batch_pull_time_sec: float
    batch_comp_time_sec: float
    batch_push_time_sec: float
    data_processing_rate: float = 0.0
    batch_custom_metrics: Dict[str, float] = field(default_factory=dict)
    kind: str = "batch"

    def to_dict(self) -> Dict[str, Any]:
        d = dict(self.__dict__)                  # flat record: one shallow copy per mini-batch (asdict recurses)
        d["batch_custom_metrics"] = dict(self.batch_custom_metrics)
        return d


@dataclass
class EpochMetrics:
    """Per-epoch worker metrics.  Reference: jobserver/src/main/avro/metrics.avsc (EpochMetrics)."""

    job_id: str
    worker_id: str
    epoch_idx: int
    num_batches_for_epoch: int
    num_epoch_data_instances: int
    epoch_time_sec: float
    epoch_custom_metrics: Dict[str, float] = field(default_factory=dict)
    kind: str = "epoch"

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)


@dataclass
class WorkerMetrics:
    """Driver-side record of one worker report (what the optimizer consumes).  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/metric/ETDolphinMetricReceiver.java:64-201, metrics.avsc
    (DolphinWorkerMetrics)."""

    worker_id: str
    hostname: str
    num_data_blocks: int
    total_pull_count: int
    bytes_received_get_resp: int
    batch: Optional[Dict[str, Any]] = None
    epoch: Optional[Dict[str, Any]] = None


@dataclass
class ServerMetrics:
    """Driver-side record of one server report.  Reference: metrics.avsc (ServerMetrics),
    ETDolphinMetricReceiver.java:64-201."""

    server_id: str
    hostname: str
    num_model_blocks: int
    total_pull_processed: int = 0
    total_push_processed: int = 0


class DolphinMetricManager:
    """Per-job store of validated worker / server metrics (feeds the optimizer and dashboard)."""

    def __init__(self, num_initial_to_skip: int = 0) -> None:
        self._lock = threading.Lock()
        self.worker_batch: Dict[str, List[Dict[str, Any]]] = {}
        self.worker_epoch: Dict[str, List[Dict[str, Any]]] = {}
        self.server: Dict[str, List[ServerMetrics]] = {}
        self._enabled = False
        self._skip = num_initial_to_skip
        self._seen: Dict[str, int] = {}
        self._skip_next_after_reconfig: Dict[str, bool] = {}
        self.expected_worker_blocks: Optional[Dict[str, int]] = None
        self.expected_server_blocks: Optional[Dict[str, int]] = None
        self.sinks: List[Any] = []     # dashboard connector etc.

    def start_metric_collection(self) -> None:
        with self._lock:
            self._enabled = True

    def stop_metric_collection(self) -> None:
        with self._lock:
            self._enabled = False

    def load_metric_validation_info(self, worker_blocks: Dict[str, int], server_blocks: Dict[str, int]) -> None:
        with self._lock:
            self.expected_worker_blocks, self.expected_server_blocks = dict(worker_blocks), dict(server_blocks)
            for w in worker_blocks:
                self._skip_next_after_reconfig[w] = True

    def store_worker_metric(self, worker_id: str, num_data_blocks: int, metric: Dict[str, Any]) -> bool:
        for s in self.sinks:
            try:
                s.post_metric(worker_id, metric)
            except Exception:  # noqa: BLE001
                pass
        with self._lock:
            if not self._enabled:
                return False
            if metric.get("kind") == "epoch":
                self.worker_epoch.setdefault(worker_id, []).append(metric)
                return True
            n = self._seen.get(worker_id, 0) + 1
            self._seen[worker_id] = n
            if n <= self._skip:
                return False
            if self._skip_next_after_reconfig.pop(worker_id, False) and self.expected_worker_blocks is not None:
                return False
            if self.expected_worker_blocks is not None and \
                    self.expected_worker_blocks.get(worker_id, num_data_blocks) != num_data_blocks:
                return False
            self.worker_batch.setdefault(worker_id, []).append(metric)
            return True

    def store_server_metric(self, m: ServerMetrics) -> bool:
        with self._lock:
            if not self._enabled:
                return False
            if self.expected_server_blocks is not None and \
                    self.expected_server_blocks.get(m.server_id, m.num_model_blocks) != m.num_model_blocks:
                return False
            self.server.setdefault(m.server_id, []).append(m)
            return True

    def get_worker_batch_metrics(self) -> Dict[str, List[Dict[str, Any]]]:
        with self._lock:
            return {k: list(v) for k, v in self.worker_batch.items()}

    def get_worker_epoch_metrics(self) -> Dict[str, List[Dict[str, Any]]]:
        with self._lock:
            return {k: list(v) for k, v in self.worker_epoch.items()}

    def get_server_metrics(self) -> Dict[str, List[ServerMetrics]]:
        with self._lock:
            return {k: list(v) for k, v in self.server.items()}

    def clear_worker_metrics(self) -> None:
        with self._lock:
            self.worker_batch.clear()
            self._seen.clear()

    def clear_server_metrics(self) -> None:
        with self._lock:
            self.server.clear()
