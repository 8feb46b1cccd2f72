"""GPU-resident NMF trainer (Dolphin Trainer SPI over ops/csrc/nmf_ops.cu + the table kernels)."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from ..dolphin.worker import Trainer
from .nmf import batch_arrays


class NMFDeviceTrainer(Trainer):
    """Dolphin Trainer over NMFDeviceModel (fused pull / compute / push on one stream)."""

    fused_step = True      # pull / compute / push are back-to-back kernels on one stream

    def __init__(self, params: Dict[str, Any], model_table, local_model_table=None, input_table=None,
                 context=None, **_: Any) -> None:
        self._params, self._context = params, context
        import torch
        from ..models.nmf import NMFDeviceModel
        self._torch = torch
        self.table, self.comps = model_table, model_table.components
        self.dtable = model_table.device_table
        self.local_table = local_model_table
        self.rank = int(params["rank"])
        # worker-local L lives in HBM: one row per local input record, in key order
        self._input_rows = list(input_table.get_local_tablet().get_data_iterator())
        self._keys = sorted(k for k, _ in self._input_rows)
        self._row_of = {k: i for i, k in enumerate(self._keys)}
        self.model = NMFDeviceModel(self.dtable, self.rank, max(1, len(self._keys)), float(params["step_size"]),
                                    float(params.get("lambda", 0.0)), float(params.get("max_val", 1e6)),
                                    float(params.get("init_min", 0.0)), float(params.get("init_max", 1.0)), seed=23)
        self._cache: Dict[Any, Any] = {}
        self._batch = None

    def _to_batch(self, mini_batch):
        from ..models.nmf import build_nmf_batch
        key = (mini_batch[0][0], len(mini_batch))
        b = self._cache.get(key)
        if b is None:
            keys, ptr, cols, vals = batch_arrays(mini_batch)
            rows = np.asarray([self._row_of[k] for k in keys], dtype=np.int32)
            b = build_nmf_batch(ptr, cols, vals, rows).to(self.dtable.device)
            self._cache[key] = b
        return b

    def set_mini_batch_data(self, mini_batch) -> None:
        self._batch = self._to_batch(mini_batch)

    def pull_model(self) -> None:
        self.comps.gate.acquire_read()
        self.model.loss.zero_()

    def local_compute(self) -> None:
        pass

    def push_update(self) -> None:
        try:
            self.model.step_batch(self._batch)      # pull + fused compute + push in stream order
            self._torch.cuda.current_stream(self.dtable.device).synchronize()
        finally:
            self.comps.gate.release_read()

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        data = list(input_data)
        return {"loss": self.model.evaluate(self._to_batch(data))} if data else {"loss": 0.0}

    def summary(self) -> Dict[str, float]:
        return {"last_batch_loss": float(self.model.loss[0])}

    def on_worker_stopped(self) -> None:
        self.cleanup()          # publish the HBM-resident local model before its table blocks migrate

    def cleanup(self) -> None:
        # publish the local model into the local-model table (checkpoint / offline evaluation)
        if self.local_table is not None and self._keys:
            L = self.model.L.cpu().numpy()
            self.local_table.multi_put(self._keys, list(L)).get(timeout=600)
            if self._params.get("print_mat"):
                from .nmf import store_matrices
                row_of = {k: rec[0] for k, rec in self._input_rows}
                store_matrices(self._params, self._context, [row_of.get(k, k) for k in self._keys], L[: len(self._keys)])
