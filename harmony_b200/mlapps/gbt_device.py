"""GPU GBT trainer: residuals, tree growth and evaluation run in ops/csrc/k12_ops.cu; the forest stays in the
(host, append-only) model table exactly as in the reference (GBTETModelUpdateFunction.java:40-48)."""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

from .gbt import GBTree, GBTTrainer


class GBTDeviceTrainer(GBTTrainer):
    """GBTTrainer whose residuals, split search, routing and evaluation run in ops/csrc/k12_ops.cu."""

    def __init__(self, params: Dict[str, Any], model_accessor, context=None, **kw: Any) -> None:
        super().__init__(params, model_accessor, **kw)
        import torch
        from ..models.gbt import DeviceTreeBuilder
        self._torch = torch
        ctx = getattr(getattr(context, "executor", None), "device_ctx", None)
        self.device = ctx.torch_device if ctx is not None else torch.device("cuda", torch.cuda.current_device())
        self.builder = DeviceTreeBuilder(self.device, self.num_features, self.types, self.max_depth, self.leaf_min,
                                         self.lam, self.gamma)

    def set_mini_batch_data(self, mini_batch) -> None:
        X, y = self._dense(mini_batch)
        self._X, self._y = X, y
        self._Xd = self._torch.from_numpy(X).to(self.device)
        self._yd = self._torch.from_numpy(y).to(self.device)
        self._order = self.builder.presort(self._Xd)

    def _predict_dev(self, forest: List[GBTree], Xd):
        return self.builder.predict(Xd, self.builder.pack_forest(forest)) if forest else \
            self._torch.zeros(Xd.shape[0], dtype=self._torch.float64, device=self.device)

    def _residual_dev(self, label: int):
        pred = self._predict_dev(self._forests[label], self._Xd)
        if self.ytype == 0:
            return self._yd - pred
        return ((self._yd - label).abs() < 1e-9).to(self._torch.float64) - pred

    def _residual(self, label: int) -> np.ndarray:
        return self._residual_dev(label).cpu().numpy()

    def _predict(self, forest: List[GBTree], X: np.ndarray) -> np.ndarray:
        Xd = self._torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(self.device)
        return self._predict_dev(forest, Xd).cpu().numpy()

    def local_compute(self) -> None:
        self._trees = []
        for lb in self.labels:
            t = self.builder.build(self._Xd, self._order, self._residual_dev(lb).contiguous(), self.step)
            self._trees.append((lb, self.builder.to_host_tree(t)))
