"""Lasso (L1-regularised linear regression) by cyclic coordinate descent with soft-thresholding.

Reference: dolphin/mlapps/lasso/LassoTrainer.java:170-224 (coordinate descent over the pulled full
model; features whose weight is ~0 are skipped), :292-322 (push step_size * (new - old) per
partition, soft threshold), LassoETParser.java:49-75 (`y idx:val ...`).  The shipped server update
function ignores the delta and re-randomises the vector (reference bug);
here the server does what the trainer expects: w += delta.
Model: key = partition p -> float[features_per_partition] (dense table, one row per partition).
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

from ..common.dataloader import DataParser, register_parser
from ..dolphin.conf import ETDolphinConfiguration
from ..dolphin.worker import Trainer
from ..et.config import TableConfiguration
from ..utils.flags import Configuration

CLOSE_TO_ZERO = 1e-9


class LassoDataParser(DataParser):
    """`y idx:val ...` -> (idx, val, y).  Reference: dolphin/mlapps/lasso/LassoETParser.java:49-75."""

    index_bound_param = "features"

    def record_indices(self, record):
        return record[0]

    native_mode = 0

    def from_arrays(self, head, ptr, idx, val):
        return [(idx[ptr[i]:ptr[i + 1]], val[ptr[i]:ptr[i + 1]], float(head[i])) for i in range(len(head))]

    def parse_line(self, line: str):
        parts = line.split()
        y = float(parts[0])
        idx = np.asarray([int(t.split(":")[0]) for t in parts[1:]], dtype=np.int32)
        val = np.asarray([float(t.split(":")[1]) for t in parts[1:]], dtype=np.float32)
        return idx, val, y


register_parser("lasso", LassoDataParser)


def sthresh(x: float, lam: float, col_norm: float) -> float:
    t = lam / col_norm
    if abs(x) <= t:
        return 0.0
    return x - t if x >= 0 else x + t


class LassoTrainer(Trainer):
    """Cyclic coordinate descent with soft-thresholding over the pulled model; pushes step * (new - old) per
    partition.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/mlapps/lasso/LassoTrainer.java:170-224,292-322."""

    def __init__(self, params: Dict[str, Any], model_accessor, **_: Any) -> None:
        self.acc = model_accessor
        self.num_features = int(params["features"])
        self.fpp = int(params.get("features_per_partition", 0) or self.num_features)
        if self.num_features % self.fpp:
            raise ValueError("Uneven model partitions")
        self.num_partitions = self.num_features // self.fpp
        self.step = float(params["step_size"])
        self.lam = float(params.get("lambda", 0.0))
        self.decay_rate = float(params.get("decay_rate", 0.9))
        self.decay_period = int(params.get("decay_period", 5))
        self.keys = np.arange(self.num_partitions, dtype=np.int64)

    def _dense(self, batch):
        X = np.zeros((len(batch), self.num_features), dtype=np.float64)
        y = np.zeros(len(batch), dtype=np.float64)
        for i, (_, (idx, val, yy)) in enumerate(batch):
            X[i, idx] = val
            y[i] = yy
        return X, y

    def set_mini_batch_data(self, mini_batch) -> None:
        self._X, self._y = self._dense(mini_batch)

    def pull_model(self) -> None:
        self._old = np.asarray(self.acc.pull_many(self.keys), dtype=np.float64).reshape(-1)
        self._new = self._old.copy()

    def local_compute(self) -> None:
        X, y, w = self._X, self._y, self._new
        pre = X @ w
        for j in range(self.num_features):
            if abs(w[j]) < CLOSE_TO_ZERO:
                continue
            col = X[:, j]
            norm = float(col @ col)
            if norm < CLOSE_TO_ZERO:
                continue
            pre -= col * w[j]
            w[j] = sthresh(float(col @ (y - pre)) / norm, self.lam, norm)
            pre += col * w[j]

    def push_update(self) -> None:
        grad = (self.step * (self._new - self._old)).astype(np.float32).reshape(self.num_partitions, self.fpp)
        self.acc.push_many(self.keys, grad)

    def on_epoch_finished(self, epoch_idx: int) -> None:
        if self.decay_rate != 1 and (epoch_idx + 1) % self.decay_period == 0:
            self.step *= self.decay_rate

    def _loss(self, data) -> float:
        if not data:
            return 0.0
        X, y = self._dense(data)
        w = np.asarray(self.acc.pull_many(self.keys), dtype=np.float64).reshape(-1)
        r = X @ w - y
        return float((r * r).sum() / (2 * len(y)) + self.lam * np.abs(w).sum())

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        return {"training_loss": self._loss(list(input_data)), "test_loss": self._loss([(None, t) for t in test_data])}

    def summary(self) -> Dict[str, float]:
        r = self._X @ self._new - self._y
        return {"last_batch_loss": float((r * r).mean())}


def model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    f = int(conf["features"])
    fpp = int(conf.get("features_per_partition", 0) or f)
    backend = conf.get("backend", "host")
    return TableConfiguration(table_id=table_id, dense=True, capacity=f // fpp, row_dim=fpp,
                              backend=backend if backend in ("host", "device") else "host",
                              update_function={"name": "add", "params": {"init": "gaussian", "mean": 0.0,
                                                                         "std": float(conf.get("model_gaussian", 0.001)),
                                                                         "seed": 0}})


APP = ETDolphinConfiguration(app_id="Lasso", trainer_class="harmony_b200.mlapps.lasso:LassoTrainer",
                             input_parser="lasso", model_table_conf=model_table_conf,
                             device_trainer_class="harmony_b200.mlapps.lasso_device:LassoDeviceTrainer")
APP_PATH = "harmony_b200.mlapps.lasso:APP"


def main_submit(argv: List[str]) -> int:
    from ..dolphin.launcher import main_submit as go
    return go(APP_PATH, APP, argv)


def main_run(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run as go
    return go(APP, argv)
