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
        self._input = input_table
        # per-epoch write-back of the HBM-resident local model is only needed when plans can move worker blocks
        self._elastic = "EmptyPlanOptimizer" not in str(params.get("optimizer", "EmptyPlanOptimizer"))
        self._input_rows = list(input_table.get_local_tablet().get_data_iterator())
        self._keys = sorted(k for k, _ in self._input_rows)
        self._row_of = {k: i for i, k in enumerate(self._keys)}
        self.model = NMFDeviceModel(self.dtable, self.rank, max(1, len(self._keys)), float(params["step_size"]),
                                    float(params.get("lambda", 0.0)), float(params.get("max_val", 1e6)),
                                    float(params.get("init_min", 0.0)), float(params.get("init_max", 1.0)), seed=23)
        self.model._used = len(self._keys)
        if local_model_table is not None and self._keys and \
                local_model_table.get_local_tablet().get_num_data_items() > 0:
            # the local-model table already holds rows here: this worker was started by a plan on blocks that
            # migrated in (or on a restored checkpoint) -- they win over the random initialisation
            rows = np.asarray(local_model_table.multi_get_or_init(self._keys).get(timeout=600), dtype=np.float32)
            self.model.L[: len(self._keys)].copy_(self._torch.from_numpy(rows[:, : self.rank]))
        self._cache: Dict[Any, Any] = {}
        self._batch = None

    # ---- worker-side elasticity (ETTrainingDataProvider.java:54-108, PlanCompiler.java:304-380) ----
    # The local model (one L row per input record) lives in HBM; the local-model TABLE travels with the input blocks
    # when a plan adds / deletes / switches workers.  So: L is written back to the table at every epoch end (and when
    # this worker is stopped), and records that show up in a mini-batch without an HBM row -- their blocks migrated
    # in -- get their rows from the table (lazily initialised there if the record is new), appended to L.
    def _adopt(self, keys) -> None:
        new = [k for k in keys if k not in self._row_of]
        if not new:
            return
        if self.local_table is not None:
            rows = np.asarray(self.local_table.multi_get_or_init(new).get(timeout=600), dtype=np.float32)
        else:
            rows = np.random.default_rng(len(self._row_of)).random((len(new), self.rank), dtype=np.float32)
        first = self.model.append_rows(self._torch.from_numpy(rows[:, : self.rank]))
        for i, k in enumerate(new):
            self._row_of[k] = first + i
        self._keys.extend(new)

    def _write_back(self) -> None:
        """Publish the L rows of the records that are still local (the others left with their blocks)."""
        if self.local_table is None or not self._keys:
            return
        here = {k for k, _ in self._input.get_local_tablet().get_data_iterator()} if self._input is not None \
            else set(self._keys)
        keys = [k for k in self._keys if k in here]
        if not keys:
            return
        L = self.model.L.cpu().numpy()
        self.local_table.multi_put(keys, [L[self._row_of[k]] for k in keys]).get(timeout=600)

    def on_epoch_finished(self, epoch_idx: int) -> None:
        super().on_epoch_finished(epoch_idx)
        if self._elastic:
            self._write_back()        # worker blocks can only move under an optimizer's plans

    def _to_batch(self, mini_batch):
        from ..models.nmf import build_nmf_batch
        key = (mini_batch[0][0], len(mini_batch))
        b = self._cache.get(key)
        if b is None:
            self._adopt([k for k, _ in mini_batch])
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
        finally:
            self.comps.gate.release_read()

    def flush(self) -> None:
        """Wait for the kernels of the mini-batches issued so far (called once per lease / epoch by the worker loop)."""
        self._torch.cuda.current_stream(self.dtable.device).synchronize()

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
            self._write_back()
            if self._params.get("print_mat"):
                from .nmf import store_matrices
                L = self.model.L.cpu().numpy()
                row_of = {k: rec[0] for k, rec in self._input_rows}
                store_matrices(self._params, self._context, [row_of.get(k, k) for k in self._keys],
                               L[[self._row_of[k] for k in self._keys]])
