"""GPU Lasso trainer: pull + cyclic coordinate descent + push are ONE kernel (ops/csrc/k12_ops.cu:lasso_cd_kernel)
reading and updating the model rows through the table's pointer table (LassoTrainer.java:170-224,292-322)."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from ..dolphin.worker import Trainer


class LassoDeviceTrainer(Trainer):
    """Dolphin Trainer over lasso_cd_kernel: pull + coordinate descent + push in one launch."""

    fused_step = True

    def __init__(self, params: Dict[str, Any], model_table, **_: Any) -> None:
        import torch
        from ..ops import native
        self._torch, self._n = torch, native()
        self.table, self.comps = model_table, model_table.components
        self.dtable = model_table.device_table
        self.device = self.dtable.device
        self.num_features = int(params["features"])
        self.fpp = int(params.get("features_per_partition", 0) or self.num_features)
        if self.num_features % self.fpp:
            raise ValueError("Uneven model partitions")
        self.step = float(params["step_size"])
        self.lam = float(params.get("lambda", 0.0))
        self.decay_rate = float(params.get("decay_rate", 0.9))
        self.decay_period = int(params.get("decay_period", 5))
        self.w_new = torch.zeros(self.num_features, dtype=torch.float32, device=self.device)
        self.w_old = torch.zeros_like(self.w_new)
        self.stats = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._cache: Dict[Any, Any] = {}

    def _to_batch(self, batch):
        key = (batch[0][0], len(batch))
        hit = self._cache.get(key)
        if hit is None:
            Xt = np.zeros((self.num_features, len(batch)), dtype=np.float32)      # column-major for the CD sweeps
            y = np.zeros(len(batch), dtype=np.float32)
            for i, (_, (idx, val, yy)) in enumerate(batch):
                Xt[idx, i] = val
                y[i] = yy
            hit = (self._torch.from_numpy(Xt).to(self.device), self._torch.from_numpy(y).to(self.device))
            if key[0] is not None:
                self._cache[key] = hit
        return hit

    def set_mini_batch_data(self, mini_batch) -> None:
        self._Xt, self._y = self._to_batch(mini_batch)

    def pull_model(self) -> None:
        self.comps.gate.acquire_read()

    def local_compute(self) -> None:
        pass

    def push_update(self) -> None:
        try:
            self._n.lasso_cd(*self.dtable.view(), self._Xt, self._y, self.lam, self.step, self.w_new, self.w_old, self.stats)
            self._torch.cuda.current_stream(self.device).synchronize()
        finally:
            self.comps.gate.release_read()

    def on_epoch_finished(self, epoch_idx: int) -> None:
        if self.decay_rate != 1 and (epoch_idx + 1) % self.decay_period == 0:
            self.step *= self.decay_rate

    def _loss(self, data) -> float:
        if not data:
            return 0.0
        Xt, y = self._to_batch(data)
        keys = self._torch.arange(self.num_features // self.fpp, dtype=self._torch.int64, device=self.device)
        w = self.dtable.pull(keys)[:, : self.fpp].reshape(-1).double()
        r = Xt.double().t() @ w - y.double()
        return float((r * r).sum() / (2 * len(y)) + self.lam * w.abs().sum())

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        return {"training_loss": self._loss(list(input_data)), "test_loss": self._loss([(None, t) for t in test_data])}

    def summary(self) -> Dict[str, float]:
        return {"last_batch_loss": float(self.stats[0])}
