"""Multinomial logistic regression (softmax regression) app.

Reference: dolphin/mlapps/mlr/MLRTrainer.java:152-160,185-187,190-300,303-316,374-398,475-505,
MLRETModelUpdateFunction.java:51-62 (init N(0, model_gaussian); w += delta), MLRETDataParser.java:52-79
(`label idx:val idx:val ...`, `#` comments), MLRJob.java:36-47, MLRParameters.java:26 (-classes).

B200 layout decision: the model table is FEATURE-major -- key f in [0, features) -> float[classes].
The reference keys the model by (class, feature-partition) and pulls/pushes the whole model every
mini-batch; with sparse instances the same mathematics only needs the rows of the features that
occur in the mini-batch, and a feature-major row is one contiguous NVLink-friendly read / reduce.
`-features_per_partition` is accepted for CLI compatibility (it only shaped the reference's keys).
With -lambda != 0 the host trainer falls back to the reference's dense pull/push (every row),
because the L2 term touches every weight; the device trainer applies the L2 term at the owners.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..common.dataloader import DataParser, register_parser
from ..dolphin import params as P
from ..dolphin.conf import ETDolphinConfiguration
from ..dolphin.worker import Trainer
from ..et.config import TableConfiguration
from ..utils.flags import Configuration

LOG = logging.getLogger(__name__)
ZERO_THRESHOLD = 1e-7


class MLRDataParser(DataParser):
    """`label idx:val idx:val ...` -> (idx int32[], val float32[], label)."""
    index_bound_param = "features"

    def record_indices(self, record):
        return record[0]

    native_mode = 0

    def from_arrays(self, head, ptr, idx, val):
        labels = head.astype(np.int64)
        nc = self.params.get("classes")
        bad = (labels < 0) | ((labels >= int(nc)) if nc else False)
        if np.any(bad):
            raise ValueError(f"label {int(labels[np.argmax(bad)])} out of range")
        return [(idx[ptr[i]:ptr[i + 1]], val[ptr[i]:ptr[i + 1]], int(labels[i])) for i in range(len(labels))]

    def parse_line(self, line: str):
        parts = line.split()
        label = int(float(parts[0]))
        nc = self.params.get("classes")
        if label < 0 or (nc and label >= int(nc)):   # (the reference accepts label == classes: App. E.5)
            raise ValueError(f"label {label} out of range")
        idx = np.empty(len(parts) - 1, dtype=np.int32)
        val = np.empty(len(parts) - 1, dtype=np.float32)
        for i, tok in enumerate(parts[1:]):
            a, _, b = tok.partition(":")
            idx[i], val[i] = int(a), float(b)
        return idx, val, label


register_parser("mlr", MLRDataParser)


def batch_to_csr(batch: Sequence[Tuple[Any, Any]]):
    n = len(batch)
    ptr = np.zeros(n + 1, dtype=np.int64)
    for i, (_, (idx, _v, _l)) in enumerate(batch):
        ptr[i + 1] = ptr[i] + len(idx)
    col = np.concatenate([b[1][0] for b in batch]) if n else np.zeros(0, dtype=np.int32)
    val = np.concatenate([b[1][1] for b in batch]) if n else np.zeros(0, dtype=np.float32)
    labels = np.fromiter((b[1][2] for b in batch), dtype=np.int64, count=n)
    return ptr, col.astype(np.int64), val.astype(np.float32), labels


def softmax_clamped(logits: np.ndarray) -> np.ndarray:
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    p = e / e.sum(axis=1, keepdims=True)
    return np.clip(p, ZERO_THRESHOLD, 1.0 - ZERO_THRESHOLD)


class MLRTrainer(Trainer):
    """Host MLR trainer on the feature-major model: sparse logits, clamped softmax, gradient rows, push; loss /
    accuracy evaluation.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/mlapps/mlr/MLRTrainer.java:190-300,374-398,404-470,475-505."""

    def __init__(self, params: Dict[str, Any], model_accessor, **_: Any) -> None:
        self.acc = model_accessor
        self.num_classes = int(params["classes"])
        self.num_features = int(params["features"])
        self.step = float(params["step_size"])
        self.lam = float(params.get("lambda", 0.0))
        self.decay_rate = float(params.get("decay_rate", 0.9))
        self.decay_period = int(params.get("decay_period", 5))
        if not (0 < self.decay_rate <= 1):
            raise ValueError("decay_rate must be larger than 0 and less than or equal to 1")
        if self.decay_period <= 0:
            raise ValueError("decay_period must be a positive value")
        self._batch = None
        self._keys: Optional[np.ndarray] = None
        self._W: Optional[np.ndarray] = None
        self._delta: Optional[np.ndarray] = None
        self.last_loss = 0.0

    # ---- SPI ----
    def set_mini_batch_data(self, mini_batch) -> None:
        self._batch = batch_to_csr(mini_batch)

    def _pull_keys(self, col: np.ndarray) -> np.ndarray:
        if self.lam != 0.0:
            return np.arange(self.num_features, dtype=np.int64)     # dense pull (L2 touches every row)
        return np.unique(col)

    def pull_model(self) -> None:
        _, col, _, _ = self._batch
        self._keys = self._pull_keys(col)
        self._W = np.asarray(self.acc.pull_many(self._keys), dtype=np.float32)

    def _forward(self, ptr, col, val, keys, W):
        n = len(ptr) - 1
        pos = np.searchsorted(keys, col)
        rows = np.repeat(np.arange(n), np.diff(ptr))
        logits = np.zeros((n, self.num_classes), dtype=np.float32)
        np.add.at(logits, rows, val[:, None] * W[pos])
        return rows, pos, softmax_clamped(logits)

    def local_compute(self) -> None:
        ptr, col, val, labels = self._batch
        rows, pos, p = self._forward(ptr, col, val, self._keys, self._W)
        n = len(labels)
        self.last_loss = float(-np.log(p[np.arange(n), labels]).sum())
        err = p
        err[np.arange(n), labels] -= 1.0
        delta = np.zeros_like(self._W)
        np.add.at(delta, pos, (-self.step) * val[:, None] * err[rows])
        if self.lam != 0.0:
            delta -= self.step * self.lam * n * self._W
        self._delta = delta

    def push_update(self) -> None:
        self.acc.push_many(self._keys, self._delta)
        self._W = self._delta = self._batch = None

    def on_epoch_finished(self, epoch_idx: int) -> None:
        if self.decay_rate != 1 and (epoch_idx + 1) % self.decay_period == 0:
            prev = self.step
            self.step *= self.decay_rate
            LOG.info("%d epochs passed. Step size decays from %g to %g", self.decay_period, prev, self.step)

    def summary(self) -> Dict[str, float]:
        return {"last_batch_loss": self.last_loss}

    def _loss_acc(self, data) -> Tuple[float, float, float]:
        if not data:
            return 0.0, 0.0, 0.0
        ptr, col, val, labels = batch_to_csr(data)
        keys = np.unique(col)
        W = np.asarray(self.acc.pull_many(keys), dtype=np.float32)
        _, _, p = self._forward(ptr, col, val, keys, W)
        n = len(labels)
        loss = float(-np.log(p[np.arange(n), labels]).sum())
        acc = float((p.argmax(axis=1) == labels).mean())
        reg = 0.0
        if self.lam != 0.0:
            Wall = np.asarray(self.acc.pull_many(np.arange(self.num_features, dtype=np.int64)), dtype=np.float64)
            reg = float((Wall ** 2).sum() * self.lam / 2 / self.num_classes)
        return loss, reg, acc

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        l, r, a = self._loss_acc(list(input_data))
        tl, tr, ta = self._loss_acc([(None, t) for t in test_data])
        return {"loss": l, "regLoss": r, "avgAccuracy": a, "testLoss": tl, "testRegLoss": tr, "testAvgAccuracy": ta}


def model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    features, classes = int(conf["features"]), int(conf["classes"])
    fpp = int(conf.get("features_per_partition", 0) or 0)
    if fpp and features % fpp != 0:
        raise ValueError("Uneven model partitions: features must be divisible by features_per_partition")
    dense_input = bool(conf.get("dense_input", False))
    tc = TableConfiguration(
        table_id=table_id, dense=True, capacity=features, row_dim=classes, dtype="float32",
        update_function={"name": "add", "params": {"init": "gaussian", "mean": 0.0,
                                                   "std": float(conf.get("model_gaussian", 0.001)), "seed": 0}},
        backend=conf.get("backend", "host") if conf.get("backend", "host") in ("host", "device") else "host")
    tc.user_params["dense_input"] = dense_input
    return tc


APP = ETDolphinConfiguration(
    app_id="MLR", trainer_class="harmony_b200.mlapps.mlr:MLRTrainer", input_parser="mlr",
    model_table_conf=model_table_conf, parameters=P.MLR,
    device_trainer_class="harmony_b200.mlapps.mlr_device:MLRDeviceTrainer")
APP_PATH = "harmony_b200.mlapps.mlr:APP"


def main_submit(argv: List[str]) -> int:
    from ..dolphin.launcher import main_submit as go
    return go(APP_PATH, APP, argv)


def main_run(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run as go
    return go(APP, argv)
